// so_desc.cu -- integer descriptor kernels on bit-packed occupancies (sm_100a).
//
// Replaces the NetworkX passes of the reference:
//   Ni-neighbour count pass            OML/LSC_cat.py:186-196, OML/NH3_cat.py:188-198
//   LSC site typing                    OML/LSC_cat.py:176-220
//   NH3 site typing (19 types)         OML/NH3_cat.py:169-396, export map :421
// The VF2 sub-graph matches collapse to closed-form local rules on the layer-3 bits (checked
// against a literal NetworkX evaluation in tests/test_oracle_types.py).  Each chain's bit-plane is staged in
// shared memory; histograms are reduced with warp ballots + popc.
#include "so_internal.cuh"
#include "so_types.cuh"
#include <stdlib.h>


// ---- pack / unpack / randomize ------------------------------------------------------------------------
__global__ void pack_kernel(const uint8_t *__restrict__ occ, uint32_t *__restrict__ bits, long long n, int N, int W,
                            unsigned long long *flag) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n * W) return;
  long long chain = t / W;
  int w = (int)(t - chain * W);
  const uint8_t *src = occ + chain * N + w * 32;
  int lim = min(32, N - w * 32);
  uint32_t word = 0;
  bool bad = false;
  for (int b = 0; b < lim; ++b) {
    uint8_t v = src[b];
    bad |= (v > 1);
    word |= (uint32_t)(v & 1) << b;
  }
  bits[t] = word;
  if (bad) atomicOr(flag, 1ULL);
}

__global__ void unpack_kernel(const uint32_t *__restrict__ bits, uint8_t *__restrict__ occ, long long n, int N, int W) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n * (long long)N) return;
  long long chain = t / N;
  int v = (int)(t - chain * N);
  occ[t] = (bits[chain * W + (v >> 5)] >> (v & 31)) & 1;
}

// Bernoulli(p) bits: word w of chain c uses Philox counters (w, i, c_lo, c_hi), i = 0..7 (32 draws);
// p per chain = coverage, or a uniform draw from counter (0xffffffff, 0, c_lo, c_hi) when coverage < 0.
__global__ void randomize_kernel(uint32_t *bits, long long n, int N, int W, unsigned long long seed, double coverage) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n * W) return;
  long long chain = t / W;
  int w = (int)(t - chain * W);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c2 = (uint32_t)chain, c3 = (uint32_t)((unsigned long long)chain >> 32);
  uint32_t o[4];
  double p = coverage;
  if (coverage < 0) {
    philox4x32_10(0xffffffffu, 0u, c2, c3, k0, k1, o);
    p = (double)o[0] * (1.0 / 4294967296.0);
  }
  // threshold on 32-bit draws: bit = draw < p * 2^32
  double thr_d = p * 4294967296.0;
  unsigned long long thr = thr_d >= 4294967296.0 ? 4294967296ULL : (unsigned long long)thr_d;
  uint32_t word = 0;
  for (int i = 0; i < 8; ++i) {
    philox4x32_10((uint32_t)w, (uint32_t)i, c2, c3, k0, k1, o);
#pragma unroll
    for (int j = 0; j < 4; ++j) word |= (uint32_t)((unsigned long long)o[j] < thr) << (i * 4 + j);
  }
  int lim = N - w * 32;
  if (lim < 32) word &= (lim <= 0) ? 0u : ((1u << lim) - 1u);
  bits[t] = word;
}

// ---- seeded island structures (drivers/generate_init_structures.py:26-116) ---------------------------------------------
// Structure i of n_strucs: np.random.seed(i); n_tops = int(3 + 15*i/(n_strucs+1)) sites drawn with
// np.random.choice(range(N), size=n_tops, replace=False), a 19-site hexagonal island of Ni around each; then n_mutate
// sites drawn the same way are toggled.  NumPy's legacy stream is restated exactly: MT19937 seeded by init_genrand,
// choice(replace=False) = the first `size` entries of permutation(N) = Fisher-Yates from the top with
// random_interval (smallest bit mask >= i, 32-bit draws, rejection).  One thread per structure; the 624-word generator
// state lives in local memory, the permutation in a global scratch row.
struct Mt19937 {
  uint32_t mt[624];
  int pos;
  __device__ void seed(uint32_t s) {
    mt[0] = s;
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    pos = 624;
  }
  __device__ uint32_t next() {
    if (pos >= 624) {
      for (int k = 0; k < 624; ++k) {
        uint32_t y = (mt[k] & 0x80000000u) | (mt[k + 1 == 624 ? 0 : k + 1] & 0x7fffffffu);
        mt[k] = mt[k + 397 >= 624 ? k + 397 - 624 : k + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      pos = 0;
    }
    uint32_t y = mt[pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
  __device__ uint32_t interval(uint32_t mx) {          // numpy random_interval, max <= 0xffffffff
    if (mx == 0) return 0;
    uint32_t mask = mx;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    uint32_t v;
    while ((v = next() & mask) > mx) {}
    return v;
  }
  __device__ void permutation(int32_t *a, int n) {     // RandomState.permutation(n): arange + legacy shuffle
    for (int i = 0; i < n; ++i) a[i] = i;
    for (int i = n - 1; i >= 1; --i) {
      int j = (int)interval((uint32_t)i);
      int32_t t = a[i]; a[i] = a[j]; a[j] = t;
    }
  }
};

__device__ __constant__ int8_t island_o1[19] = {-2, -1, 0, -2, -1, 0, 1, -2, -1, 0, 1, 2, -1, 0, 1, 2, 0, 1, 2};
__device__ __constant__ int8_t island_o2[19] = {2, 2, 2, 1, 1, 1, 1, 0, 0, 0, 0, 0, -1, -1, -1, -1, -2, -2, -2};

__global__ void island_kernel(uint32_t *bits, long long n, int d, int N, int W, const int32_t *index, int n_strucs,
                              int n_mutate, int32_t *scratch) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n) return;
  uint32_t *x = bits + t * W;
  int32_t *perm = scratch + t * N;
  for (int w = 0; w < W; ++w) x[w] = 0;
  const int i = index[t];
  Mt19937 g;
  g.seed((uint32_t)i);
  const int n_tops = (int)(3.0 + 15.0 * (double)i / (double)(n_strucs + 1));
  g.permutation(perm, N);
  for (int k = 0; k < n_tops && k < N; ++k) {
    const int top = perm[k], t2 = top / d, t1 = top - t2 * d;
    for (int q = 0; q < 19; ++q) {
      int v = wrap(t2 + island_o2[q], d) * d + wrap(t1 + island_o1[q], d);
      x[v >> 5] |= 1u << (v & 31);
    }
  }
  g.permutation(perm, N);
  for (int k = 0; k < n_mutate && k < N; ++k) {
    const int v = perm[k];
    x[v >> 5] ^= 1u << (v & 31);
  }
}

// ---- full-structure descriptors: one block per chain ------------------------------------------------
__global__ void __launch_bounds__(256) descriptors_kernel(const uint32_t *__restrict__ bits, int d, int N, int W,
                                                          const int32_t *__restrict__ h5_ptr,
                                                          const int32_t *__restrict__ h5_idx, uint8_t *c6_out,
                                                          uint8_t *lsc_out, uint8_t *nh3_out, int32_t *lsc_hist,
                                                          int32_t *nh3_hist, int32_t *nh3_exp, int want_nh3) {
  extern __shared__ uint32_t s_bits[];
  __shared__ int s_hist[24];
  const long long chain = blockIdx.x;
  for (int w = threadIdx.x; w < W; w += blockDim.x) s_bits[w] = bits[chain * W + w];
  if (threadIdx.x < 24) s_hist[threadIdx.x] = 0;
  __syncthreads();
  BitView b{s_bits, d, -1};
  const int lane = threadIdx.x & 31;
  int cnt_nh3 = 0;   // lane t counts NH3 type t (t < 20)
  int cnt_lsc = 0;   // lane t counts LSC type t (t < 3)
  const int first = want_nh3 ? N : 3 * N, last = want_nh3 ? 5 * N : 4 * N;
  for (int base = first; base < last; base += blockDim.x) {
    int n = base + threadIdx.x;
    int t = 255, tl = 255;
    if (n < last) {
      int layer = n / N;
      int v = n - layer * N;
      int i2 = v / d, i1 = v - i2 * d;
      if (layer == 3) {
        L3Info s = l3_info(b, i1, i2);
        tl = lsc_type_of(s);
        if (c6_out) c6_out[chain * N + v] = (uint8_t)s.c6;
        if (lsc_out) lsc_out[chain * N + v] = (uint8_t)tl;
        if (want_nh3) t = nh3_l3_type(b, i1, i2, s, h5_ptr, h5_idx);
      } else if (layer == 4) {
        t = nh3_l4_type(b, i1, i2);
      } else if (layer == 2) {
        t = nh3_l2_type(b, i1, i2);
      } else {
        t = nh3_l1_type(b, i1, i2);
      }
      if (want_nh3 && nh3_out) nh3_out[chain * 5LL * N + n] = (uint8_t)t;
    }
    if (want_nh3) {
#pragma unroll
      for (int tt = 0; tt < 20; ++tt) {
        unsigned m = __ballot_sync(SO_FULL, t == tt);
        if (lane == tt) cnt_nh3 += __popc(m);
      }
    }
#pragma unroll
    for (int tt = 0; tt < 3; ++tt) {
      unsigned m = __ballot_sync(SO_FULL, tl == tt);
      if (lane == tt) cnt_lsc += __popc(m);
    }
  }
  if (lane < 20 && cnt_nh3) atomicAdd(&s_hist[lane], cnt_nh3);
  if (lane < 3 && cnt_lsc) atomicAdd(&s_hist[20 + lane], cnt_lsc);
  __syncthreads();
  if (threadIdx.x < 20 && want_nh3) {
    int hv = s_hist[threadIdx.x] + (threadIdx.x == 0 ? N : 0);   // layer 0 is always None
    if (nh3_hist) nh3_hist[chain * 20 + threadIdx.x] = hv;
  }
  if (threadIdx.x < 3 && lsc_hist) lsc_hist[chain * 3 + threadIdx.x] = s_hist[20 + threadIdx.x];
  if (threadIdx.x == 0 && want_nh3 && nh3_exp) {
    int e1 = s_hist[1], e2 = s_hist[3], e3 = s_hist[5], e4 = s_hist[16], e5 = s_hist[17], e6 = s_hist[18],
        e7 = s_hist[19];
    int32_t *o = nh3_exp + chain * 8;
    o[0] = 5 * N - (e1 + e2 + e3 + e4 + e5 + e6 + e7);
    o[1] = e1; o[2] = e2; o[3] = e3; o[4] = e4; o[5] = e5; o[6] = e6; o[7] = e7;
  }
}

// ---- bit-sliced histograms: 32 sites per integer operation ------------------------------------------------------
// Every typing rule is a boolean function of neighbouring occupancy bits, so whole lattice rows are processed as
// bit-vectors: a thread owns one 32-bit word of one row; neighbour access is a 1-bit rotation of a row (periodic in
// i1) or the row above / below (periodic in i2); counts of Ni among 6 ring or 3 triangle sites are bit-sliced adders;
// histograms are popc of the resulting type masks.  Three passes (layer 3 -> layers 4 and 2 -> vacancy refinements
// and layer 1) exchange intermediate planes through shared memory.  Used when only histograms are requested.
struct Planes {
  uint32_t *X, *p3, *p5, *p45, *p12, *p14, *p8, *p15, *pn2;
  int d, RW, tail;
  uint32_t tailmask;
};
__device__ __forceinline__ uint32_t pl_get(const Planes &P, const uint32_t *pl, int row, int w) { return pl[row * P.RW + w]; }
__device__ __forceinline__ uint32_t pl_valid(const Planes &P, int w) { return w == P.RW - 1 ? P.tailmask : 0xffffffffu; }
// result bit i = pl[row][i + 1 mod d]
__device__ __forceinline__ uint32_t pl_rotr1(const Planes &P, const uint32_t *pl, int row, int w) {
  const uint32_t *r = pl + row * P.RW;
  uint32_t x = r[w];
  if (w < P.RW - 1) return (x >> 1) | (r[w + 1] << 31);
  return ((x >> 1) | ((r[0] & 1u) << (P.tail - 1))) & P.tailmask;
}
// result bit i = pl[row][i - 1 mod d]
__device__ __forceinline__ uint32_t pl_rotl1(const Planes &P, const uint32_t *pl, int row, int w) {
  const uint32_t *r = pl + row * P.RW;
  uint32_t in = w > 0 ? (r[w - 1] >> 31) : ((r[P.RW - 1] >> (P.tail - 1)) & 1u);
  return ((r[w] << 1) | in) & pl_valid(P, w);
}
// NON-periodic: result bit i = pl[row][i + o] if 0 <= i + o < d else 0; rows outside [0, d) read as 0
__device__ __forceinline__ uint32_t pl_shift_np(const Planes &P, const uint32_t *pl, int row, int w, int o) {
  if (row < 0 || row >= P.d) return 0u;
  const uint32_t *r = pl + row * P.RW;
  uint32_t x = r[w], out;
  if (o == 0) out = x;
  else if (o > 0) out = (x >> o) | (w + 1 < P.RW ? r[w + 1] << (32 - o) : 0u);
  else out = (x << -o) | (w > 0 ? r[w - 1] >> (32 + o) : 0u);
  return out & pl_valid(P, w);
}
__device__ __forceinline__ void add3(uint32_t a, uint32_t b, uint32_t c, uint32_t &s0, uint32_t &s1) {
  s0 = a ^ b ^ c;
  s1 = (a & b) | (c & (a ^ b));
}

__global__ void __launch_bounds__(128) descriptors_bits_kernel(const uint32_t *__restrict__ bits, int d, int N, int W,
                                                               int32_t *lsc_hist, int32_t *nh3_hist, int32_t *nh3_exp,
                                                               int want_nh3) {
  extern __shared__ uint32_t s_pl[];
  __shared__ int s_hist[24];
  const long long chain = blockIdx.x;
  Planes P;
  P.d = d; P.RW = (d + 31) / 32; P.tail = d - 32 * (P.RW - 1);
  P.tailmask = P.tail == 32 ? 0xffffffffu : ((1u << P.tail) - 1u);
  const int items = d * P.RW;
  P.X = s_pl; P.p3 = P.X + items; P.p5 = P.p3 + items; P.p45 = P.p5 + items; P.p12 = P.p45 + items;
  P.p14 = P.p12 + items; P.p8 = P.p14 + items; P.p15 = P.p8 + items; P.pn2 = P.p15 + items;
  if (threadIdx.x < 24) s_hist[threadIdx.x] = 0;
  // row-aligned copy of the packed bit-plane (site v = row*d + i1)
  const uint32_t *src = bits + chain * W;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    int row = it / P.RW, w = it - row * P.RW;
    int v0 = row * d + 32 * w, wi = v0 >> 5;
    uint32_t lo = src[wi], hi = wi + 1 < W ? src[wi + 1] : 0u;
    P.X[it] = __funnelshift_r(lo, hi, v0 & 31) & pl_valid(P, w);
  }
  __syncthreads();
  int cnt[20];
#pragma unroll
  for (int t = 0; t < 20; ++t) cnt[t] = 0;
  int lsc1 = 0, lsc2 = 0;
  // ---- pass A: layer 3, Ni sites: c6 by a bit-sliced adder, types 3 / 5 / 4 (LSC 1 / 2) ---------------------------
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    int row = it / P.RW, w = it - row * P.RW;
    int rp = row + 1 == d ? 0 : row + 1, rm = row == 0 ? d - 1 : row - 1;
    uint32_t x = P.X[it];
    uint32_t r0 = pl_rotr1(P, P.X, row, w), r1 = pl_get(P, P.X, rp, w), r2 = pl_rotl1(P, P.X, rp, w);
    uint32_t r3 = pl_rotl1(P, P.X, row, w), r4 = pl_get(P, P.X, rm, w), r5 = pl_rotr1(P, P.X, rm, w);
    uint32_t s1, c1, s2, c2, b1, b2;
    add3(r0, r1, r2, s1, c1);
    add3(r3, r4, r5, s2, c2);
    uint32_t b0 = s1 ^ s2, k = s1 & s2;
    add3(c1, c2, k, b1, b2);
    uint32_t eq6 = b2 & b1 & ~b0, eq4 = b2 & ~b1 & ~b0;
    uint32_t adj = ~(r0 | r1) | ~(r1 | r2) | ~(r2 | r3) | ~(r3 | r4) | ~(r4 | r5) | ~(r5 | r0);
    uint32_t t3 = x & eq6, t5 = x & eq4 & adj, t4 = x & ~t3 & ~t5;
    P.p3[it] = t3; P.p5[it] = t5; P.p45[it] = t4 | t5;
    lsc1 += __popc(t3); lsc2 += __popc(t5);
    cnt[3] += __popc(t3); cnt[5] += __popc(t5); cnt[4] += __popc(t4);
  }
  if (want_nh3) {
    __syncthreads();
    // ---- pass B: layer 4 (triangle under it: (0,-1),(1,-1),(0,0)) and layer 2 (over it: (1,-1),(0,0),(1,0)) -------
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
      int row = it / P.RW, w = it - row * P.RW;
      int rm = row == 0 ? d - 1 : row - 1;
      uint32_t valid = pl_valid(P, w);
      {
        uint32_t a = pl_get(P, P.X, rm, w), b = pl_rotr1(P, P.X, rm, w), c = pl_get(P, P.X, row, w), n0, n1, e0, e1;
        add3(a, b, c, n0, n1);
        add3(pl_get(P, P.p45, rm, w), pl_rotr1(P, P.p45, rm, w), pl_get(P, P.p45, row, w), e0, e1);
        uint32_t all3 = n0 & n1;
        uint32_t t1 = all3 & ~e0 & ~e1, t17 = all3 & e0 & ~e1, t16 = all3 & e1, t12 = n0 & ~n1, t14 = ~n0 & n1;
        P.p12[it] = t12; P.p14[it] = t14;
        cnt[1] += __popc(t1); cnt[17] += __popc(t17); cnt[16] += __popc(t16); cnt[12] += __popc(t12); cnt[14] += __popc(t14);
      }
      {
        uint32_t a = pl_rotr1(P, P.X, rm, w), b = pl_get(P, P.X, row, w), c = pl_rotr1(P, P.X, row, w), n0, n1, t0, t1m, e0, e1m;
        add3(a, b, c, n0, n1);
        add3(pl_rotr1(P, P.p3, rm, w), pl_get(P, P.p3, row, w), pl_rotr1(P, P.p3, row, w), t0, t1m);
        add3(pl_rotr1(P, P.p5, rm, w), pl_get(P, P.p5, row, w), pl_rotr1(P, P.p5, row, w), e0, e1m);
        uint32_t all3 = n0 & n1;
        uint32_t t18 = all3 & (~t0 & t1m) & (e0 & ~e1m), t19 = all3 & (t0 & ~t1m) & (~e0 & e1m);
        uint32_t t2 = all3 & ~t18 & ~t19, t8 = ~n0 & ~n1 & valid, t15 = ~n0 & n1, tn = n0 & ~n1;
        P.p8[it] = t8; P.p15[it] = t15; P.pn2[it] = tn;
        cnt[18] += __popc(t18); cnt[19] += __popc(t19); cnt[2] += __popc(t2); cnt[8] += __popc(t8); cnt[15] += __popc(t15);
      }
    }
    __syncthreads();
    // ---- pass C: layer-3 vacancies (11 / 10 / 9 / 6) and layer 1 (7 / 13) ------------------------------------------
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
      int row = it / P.RW, w = it - row * P.RW;
      int rp = row + 1 == d ? 0 : row + 1, rm = row == 0 ? d - 1 : row - 1;
      uint32_t valid = pl_valid(P, w);
      {
        uint32_t vac = ~P.X[it] & valid;
        // layer-4 atoms above a site: (0,0), (-1,1), (0,1)
        uint32_t has14 = pl_get(P, P.p14, row, w) | pl_rotl1(P, P.p14, rp, w) | pl_get(P, P.p14, rp, w);
        uint32_t has12 = pl_get(P, P.p12, row, w) | pl_rotl1(P, P.p12, rp, w) | pl_get(P, P.p12, rp, w);
        uint32_t t11 = vac & has14, t10 = vac & has12 & ~has14, rest = vac & ~has12 & ~has14;
        // type-15 layer-2 atoms within 2*2.77 A on RAW (non-periodic) positions: 12 index offsets (k - i)
        uint32_t near = 0;
        if (rest) {
          near |= pl_shift_np(P, P.p15, row, w, -2) | pl_shift_np(P, P.p15, row + 1, w, -2) | pl_shift_np(P, P.p15, row + 2, w, -2);
          near |= pl_shift_np(P, P.p15, row - 1, w, -1) | pl_shift_np(P, P.p15, row, w, -1) |
                  pl_shift_np(P, P.p15, row + 1, w, -1) | pl_shift_np(P, P.p15, row + 2, w, -1);
          near |= pl_shift_np(P, P.p15, row - 1, w, 0) | pl_shift_np(P, P.p15, row, w, 0) | pl_shift_np(P, P.p15, row + 1, w, 0);
          near |= pl_shift_np(P, P.p15, row - 1, w, 1) | pl_shift_np(P, P.p15, row, w, 1);
        }
        uint32_t t9 = rest & near, t6 = rest & ~near;
        cnt[11] += __popc(t11); cnt[10] += __popc(t10); cnt[9] += __popc(t9); cnt[6] += __popc(t6);
      }
      {
        // layer-2 atoms above a layer-1 atom: (-1,0), (0,-1), (0,0)
        uint32_t a = pl_rotl1(P, P.p8, row, w), b = pl_get(P, P.p8, rm, w), c = pl_get(P, P.p8, row, w), n0, n1, z0, z1;
        add3(a, b, c, n0, n1);
        add3(pl_rotl1(P, P.pn2, row, w), pl_get(P, P.pn2, rm, w), pl_get(P, P.pn2, row, w), z0, z1);
        uint32_t t7 = n0 & n1 & valid, t13 = (~n0 & n1) & (z0 & ~z1) & valid;
        cnt[7] += __popc(t7); cnt[13] += __popc(t13);
      }
    }
  }
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int t = 1; t < 20; ++t) {
    int v = __reduce_add_sync(SO_FULL, cnt[t]);
    if (lane == 0 && v) atomicAdd(&s_hist[t], v);
  }
  lsc1 = __reduce_add_sync(SO_FULL, lsc1);
  lsc2 = __reduce_add_sync(SO_FULL, lsc2);
  if (lane == 0) { if (lsc1) atomicAdd(&s_hist[21], lsc1); if (lsc2) atomicAdd(&s_hist[22], lsc2); }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (lsc_hist) {
      lsc_hist[chain * 3 + 0] = N - s_hist[21] - s_hist[22];
      lsc_hist[chain * 3 + 1] = s_hist[21];
      lsc_hist[chain * 3 + 2] = s_hist[22];
    }
    if (want_nh3) {
      int typed = 0;
      for (int t = 1; t < 20; ++t) typed += s_hist[t];
      if (nh3_hist) {
        nh3_hist[chain * 20] = 5 * N - typed;
        for (int t = 1; t < 20; ++t) nh3_hist[chain * 20 + t] = s_hist[t];
      }
      if (nh3_exp) {
        int e1 = s_hist[1], e2 = s_hist[3], e3 = s_hist[5], e4 = s_hist[16], e5 = s_hist[17], e6 = s_hist[18],
            e7 = s_hist[19];
        int32_t *o = nh3_exp + chain * 8;
        o[0] = 5 * N - (e1 + e2 + e3 + e4 + e5 + e6 + e7);
        o[1] = e1; o[2] = e2; o[3] = e3; o[4] = e4; o[5] = e5; o[6] = e6; o[7] = e7;
      }
    }
  }
}

// ---- per-flip descriptor deltas: one warp per chain, one affected node per lane -----------------------
__global__ void __launch_bounds__(256) flip_delta_kernel(const uint32_t *__restrict__ bits, long long n, int d, int N,
                                                         int W, const int32_t *__restrict__ sites,
                                                         int32_t *d_lsc, int32_t *d_nh3_exp) {
  const int lane = threadIdx.x & 31;
  long long chain = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (chain >= n) return;
  const uint32_t *w = bits + chain * W;
  const int dlt = so_flip_hist_delta(w, d, N, sites[chain], lane);
  if (lane < 8 && d_nh3_exp) d_nh3_exp[chain * 8 + lane] = dlt;
  if (lane >= 8 && lane < 11 && d_lsc) d_lsc[chain * 3 + lane - 8] = dlt;
}

// ---- launch wrappers -------------------------------------------------------------------------------------
int so_launch_pack(const uint8_t *occ_dev, uint32_t *bits_dev, int64_t n, int N, int W, unsigned long long *flag,
                   cudaStream_t st) {
  long long total = n * W;
  if (total == 0) return SO_OK;
  pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(occ_dev, bits_dev, n, N, W, flag);
  SO_CUDA(cudaGetLastError());
  return SO_OK;
}
int so_launch_unpack(const uint32_t *bits_dev, uint8_t *occ_dev, int64_t n, int N, int W, cudaStream_t st) {
  long long total = n * (long long)N;
  if (total == 0) return SO_OK;
  unpack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(bits_dev, occ_dev, n, N, W);
  SO_CUDA(cudaGetLastError());
  return SO_OK;
}
int so_launch_randomize(uint32_t *bits_dev, int64_t n, int N, int W, uint64_t seed, double coverage, cudaStream_t st) {
  long long total = n * W;
  if (total == 0) return SO_OK;
  randomize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(bits_dev, n, N, W, seed, coverage);
  SO_CUDA(cudaGetLastError());
  return SO_OK;
}
int so_launch_islands(uint32_t *bits_dev, int64_t n, int d, int N, int W, const int32_t *index_dev, int n_strucs, int n_mutate,
                      int32_t *scratch_dev, cudaStream_t st) {
  if (n == 0) return SO_OK;
  island_kernel<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(bits_dev, n, d, N, W, index_dev, n_strucs, n_mutate, scratch_dev);
  SO_CUDA(cudaGetLastError());
  return SO_OK;
}
int so_launch_descriptors(const so_lattice *lat, const uint32_t *bits_dev, int64_t n, uint8_t *c6, uint8_t *lsc_type,
                          uint8_t *nh3_type, int32_t *lsc_hist, int32_t *nh3_hist, int32_t *nh3_exp, cudaStream_t st) {
  if (n == 0) return SO_OK;
  int want_nh3 = (nh3_type || nh3_hist || nh3_exp) ? 1 : 0;
  {
    // histograms only: the bit-sliced kernel (32 sites per integer operation); SO_DESC_BITS=0 keeps the per-site kernel
    const char *env = getenv("SO_DESC_BITS");
    size_t smem = (size_t)9 * lat->d * ((lat->d + 31) / 32) * 4;
    if (!c6 && !lsc_type && !nh3_type && smem <= 96 * 1024 && !(env && env[0] == '0')) {
      SO_CUDA(cudaFuncSetAttribute(descriptors_bits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      descriptors_bits_kernel<<<(unsigned)n, 128, smem, st>>>(bits_dev, lat->d, lat->N, lat->W, lsc_hist, nh3_hist, nh3_exp,
                                                              want_nh3);
      SO_CUDA(cudaGetLastError());
      return SO_OK;
    }
  }
  int threads = lat->N >= 256 ? 256 : ((lat->N + 31) / 32) * 32;
  descriptors_kernel<<<(unsigned)n, threads, lat->W * sizeof(uint32_t), st>>>(
      bits_dev, lat->d, lat->N, lat->W, lat->h5_ptr_dev, lat->h5_idx_dev, c6, lsc_type, nh3_type, lsc_hist, nh3_hist,
      nh3_exp, want_nh3);
  SO_CUDA(cudaGetLastError());
  return SO_OK;
}
int so_launch_flip_delta(const so_lattice *lat, const uint32_t *bits_dev, int64_t n, const int32_t *sites_dev,
                         int32_t *d_lsc, int32_t *d_nh3_exp, cudaStream_t st) {
  if (n == 0) return SO_OK;
  long long threads = n * 32;
  flip_delta_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(bits_dev, n, lat->d, lat->N, lat->W, sites_dev,
                                                                       d_lsc, d_nh3_exp);
  SO_CUDA(cudaGetLastError());
  return SO_OK;
}
