// so_sa.cu -- surrogate state, per-flip delta evaluation and the Metropolis loop (sm_100a).
//
// Replaces, for n independent chains at once:
//   surrogate.eval_rate(all translations)          OML/train_surrogate.py:34-54, OML/optimize_SA.py:41-42
//   the hot loop of optimize()                     OML/optimize_SA.py:59-132
// Arithmetic: the first-layer pre-activations of every site are kept as exact int32 fixed-point
// sums hq[site][Hp]; a flip at f changes one input of each affected row r = f - offset_k, so
//   hq_r += +-W1q[k]   and   dA = sum_r sum_j w2q_j * (relu(hq'_rj) - relu(hq_rj))   (exact int64),
// which is order independent and therefore reproduced bit-for-bit by oracle/sa.py.  One warp owns one chain;
// the state streams from HBM (the kernel is bandwidth bound: 4*K*Hp*(1+accept) bytes per flip-evaluation).
#include "so_internal.cuh"
#include "so_types.cuh"
#include <math.h>
#include <stdlib.h>

// ---- score: hq, row sums, gate, objective sums; one thread per row -------------------------------------
template <int HP4>
__global__ void __launch_bounds__(256) score_kernel(const uint32_t *__restrict__ bits, int d, int N, int W,
                                                    SurView sv, int tab_in_smem, int32_t *hq, long long *hi_out,
                                                    long long *lo_out, int32_t *nact_out) {
  constexpr int HP = HP4 * 4;
  extern __shared__ __align__(16) unsigned char s_raw[];
  uint32_t *s_bits = (uint32_t *)s_raw;
  int4 *s_W = (int4 *)(s_raw + ((W * 4 + 15) / 16) * 16);
  int32_t *s_off = (int32_t *)(s_W + (tab_in_smem ? sv.K * HP4 : 0));
  __shared__ long long s_hi, s_lo;
  __shared__ int s_n;
  const long long chain = blockIdx.y;
  for (int w = threadIdx.x; w < W; w += blockDim.x) s_bits[w] = bits[chain * W + w];
  if (tab_in_smem) {
    for (int q = threadIdx.x; q < sv.K * HP4; q += blockDim.x) s_W[q] = ((const int4 *)sv.W1q)[q];
    for (int k = threadIdx.x; k < sv.K; k += blockDim.x) s_off[k] = sv.offs[k];
  }
  if (threadIdx.x == 0) { s_hi = 0; s_lo = 0; s_n = 0; }
  __syncthreads();
  const int4 *Wt = tab_in_smem ? s_W : (const int4 *)sv.W1q;
  const int32_t *offt = tab_in_smem ? s_off : sv.offs;
  BitView b{s_bits, d, -1};
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  long long my_hi = 0, my_lo = 0;
  int my_n = 0;
  if (r < N) {
    int r2 = r / d, r1 = r - r2 * d;
    int acc[HP];
#pragma unroll
    for (int j = 0; j < HP; ++j) acc[j] = __ldg(sv.b1q + j);
    for (int k = 0; k < sv.K; ++k) {
      int off = offt[k];
      int o1 = (int)(short)(off & 0xffff), o2 = off >> 16;
      int bit = b.at(r1 + o1, r2 + o2);
#pragma unroll
      for (int p = 0; p < HP4; ++p) {
        int4 w = Wt[k * HP4 + p];
        acc[4 * p + 0] += bit * w.x; acc[4 * p + 1] += bit * w.y;
        acc[4 * p + 2] += bit * w.z; acc[4 * p + 3] += bit * w.w;
      }
    }
    long long A = 0;
#pragma unroll
    for (int j = 0; j < HP; ++j) A += (long long)__ldg(sv.w2q + j) * (long long)max(acc[j], 0);
    if (hq) {
      int4 *dst = (int4 *)(hq + (chain * N + r) * (long long)HP);
#pragma unroll
      for (int p = 0; p < HP4; ++p) dst[p] = make_int4(acc[4 * p], acc[4 * p + 1], acc[4 * p + 2], acc[4 * p + 3]);
    }
    int g = sv.gated ? tree_gate(sv, b, r1, r2) : 1;
    if (g) { my_hi = A >> SO_SPLIT_BITS; my_lo = A & SO_SPLIT_MASK; my_n = 1; }
  }
  my_hi = warp_sum_ll(my_hi);
  my_lo = warp_sum_ll(my_lo);
  my_n = warp_sum_i(my_n);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd((unsigned long long *)&s_hi, (unsigned long long)my_hi);
    atomicAdd((unsigned long long *)&s_lo, (unsigned long long)my_lo);
    atomicAdd(&s_n, my_n);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd((unsigned long long *)(hi_out + chain), (unsigned long long)s_hi);
    atomicAdd((unsigned long long *)(lo_out + chain), (unsigned long long)s_lo);
    atomicAdd(nact_out + chain, s_n);
  }
}

// canonical form 0 <= lo < 2^20 after the un-ordered accumulation above
__global__ void normalize_sums_kernel(long long *hi, long long *lo, long long n) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n) return;
  long long l = lo[t];
  hi[t] += l >> SO_SPLIT_BITS;
  lo[t] = l & SO_SPLIT_MASK;
}

__global__ void objective_kernel(SurView sv, const long long *hi, const long long *lo, const int32_t *nact, long long n,
                                 double *of) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n) return;
  of[t] = so_real(sv.c, sv.b2, sv.y_scale, sv.y_mean, hi[t], lo[t], nact[t]);
}

// generic forward over explicit descriptor rows X[n_rows][K] (bytes)
template <int HP4>
__global__ void __launch_bounds__(128) eval_rows_kernel(SurView sv, long long n_rows, const uint8_t *__restrict__ X,
                                                        uint8_t *active, double *rate) {
  constexpr int HP = HP4 * 4;
  long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const uint8_t *x = X + r * sv.K;
  int acc[HP];
#pragma unroll
  for (int j = 0; j < HP; ++j) acc[j] = __ldg(sv.b1q + j);
  for (int k = 0; k < sv.K; ++k) {
    int bit = x[k] ? 1 : 0;
#pragma unroll
    for (int p = 0; p < HP4; ++p) {
      int4 w = __ldg((const int4 *)sv.W1q + k * HP4 + p);
      acc[4 * p + 0] += bit * w.x; acc[4 * p + 1] += bit * w.y;
      acc[4 * p + 2] += bit * w.z; acc[4 * p + 3] += bit * w.w;
    }
  }
  long long A = 0;
#pragma unroll
  for (int j = 0; j < HP; ++j) A += (long long)__ldg(sv.w2q + j) * (long long)max(acc[j], 0);
  int g = 1;
  if (sv.gated) {
    int node = 0;
    int f = __ldg(sv.t_feature);
    while (f >= 0) {
      node = ((double)(x[f] ? 1 : 0) <= __ldg(sv.t_thr + node)) ? __ldg(sv.t_left + node) : __ldg(sv.t_right + node);
      f = __ldg(sv.t_feature + node);
    }
    g = __ldg(sv.t_active + node);
  }
  if (active) active[r] = (uint8_t)g;
  if (rate) rate[r] = so_real(sv.c, sv.b2, sv.y_scale, sv.y_mean, A >> SO_SPLIT_BITS, A & SO_SPLIT_MASK, 1);
}

// ---- proposal / uniform draw -----------------------------------------------------------------------------
__device__ __forceinline__ void draw(const SaParams &P, long long c, long long s, int &f, float &u) {
  if (P.proposals) {
    f = __ldg(P.proposals + c * P.n_steps + s);
    u = __ldg(P.uniforms + c * P.n_steps + s);
  } else {
    unsigned long long step = (unsigned long long)(P.step0 + s), cid = (unsigned long long)(P.chain_id0 + c);
    uint32_t o[4];
    philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), (uint32_t)cid, (uint32_t)(cid >> 32), (uint32_t)P.seed,
                  (uint32_t)(P.seed >> 32), o);
    f = (int)__umulhi(o[0], (uint32_t)P.N);
    u = (float)(o[1] >> 8) * 5.9604644775390625e-08f;   // 2^-24, exact
  }
}

// Metropolis (OML/optimize_SA.py:108-116).  For T > 0 the rule is "accept iff delta > T ln u" (delta > 0
// implies it), for T <= 0 "accept iff delta > 0"; `guard` is set when delta is within tol of that threshold.
__device__ __forceinline__ bool metropolis(double delta, double T, float u, double tol, bool &guard) {
  bool acc;
  if (delta > 0.0) acc = true;
  else if (T > 0.0) acc = exp(delta / T) > (double)u;
  else acc = false;
  guard = tol > 0.0 && so_near_tie(delta, T, u, tol);
  return acc;
}

// ---- flat kernel: no gate.  One warp per chain; lane-strided over the K*HP4 int4 of the affected rows -----
template <int HP4, int MAXIT, bool TAB_SMEM>
__global__ void __launch_bounds__(256) sa_flat_kernel(SaParams P) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const SurView &sv = P.sv;
  const int Q = sv.K * HP4;
  int4 *s_W = (int4 *)s_raw;
  int4 *s_w2 = s_W + (TAB_SMEM ? Q : 0);
  int32_t *s_off = (int32_t *)(s_w2 + HP4);
  if (TAB_SMEM) {
    for (int q = threadIdx.x; q < Q; q += blockDim.x) s_W[q] = ((const int4 *)sv.W1q)[q];
    for (int k = threadIdx.x; k < sv.K; k += blockDim.x) s_off[k] = sv.offs[k];
  }
  for (int p = threadIdx.x; p < HP4; p += blockDim.x) s_w2[p] = ((const int4 *)sv.w2q)[p];
  __syncthreads();
  const int4 *Wt = TAB_SMEM ? s_W : (const int4 *)sv.W1q;
  const int32_t *offt = TAB_SMEM ? s_off : sv.offs;
  const int lane = threadIdx.x & 31;
  const int d = P.d;

  for (;;) {
    long long c = 0;
    if (lane == 0) c = (long long)atomicAdd(P.counters, 1ULL);
    c = __shfl_sync(SO_FULL, c, 0);
    if (c >= P.n_chains) break;
    uint32_t *bits = P.bits + c * P.W;
    int4 *hq = (int4 *)P.hq + c * (long long)P.N * HP4;
    long long hi = P.hi[c], lo = P.lo[c];
    const double tsc = P.tscale[c];
    unsigned n_acc = 0, n_guard = 0;

    for (long long s = 0; s < P.n_steps; ++s) {
      const double T = __dmul_rn(tsc, __ldg(P.temps + s));
      int f;
      float u;
      draw(P, c, s, f, u);
      const int f2 = f / d, f1 = f - f2 * d;
      const uint32_t word = bits[f >> 5];
      const int sgn = ((word >> (f & 31)) & 1) ? -1 : 1;
      long long acc = 0;
      int4 hn[MAXIT > 0 ? MAXIT : 1];
      int idx[MAXIT > 0 ? MAXIT : 1];
      if (MAXIT > 0) {
#pragma unroll
        for (int it = 0; it < MAXIT; ++it) {
          int q = lane + 32 * it;
          if (q < Q) {
            int k = q / HP4, part = q - k * HP4;
            int off = offt[k];
            int r1 = wrap(f1 - (int)(short)(off & 0xffff), d), r2 = wrap(f2 - (off >> 16), d);
            idx[it] = (r2 * d + r1) * HP4 + part;
            int4 h = hq[idx[it]];
            int4 w = Wt[q], w2 = s_w2[part], n;
            n.x = h.x + sgn * w.x; n.y = h.y + sgn * w.y; n.z = h.z + sgn * w.z; n.w = h.w + sgn * w.w;
            acc += (long long)w2.x * (long long)(max(n.x, 0) - max(h.x, 0));
            acc += (long long)w2.y * (long long)(max(n.y, 0) - max(h.y, 0));
            acc += (long long)w2.z * (long long)(max(n.z, 0) - max(h.z, 0));
            acc += (long long)w2.w * (long long)(max(n.w, 0) - max(h.w, 0));
            hn[it] = n;
          }
        }
      } else {
        for (int q = lane; q < Q; q += 32) {
          int k = q / HP4, part = q - k * HP4;
          int off = offt[k];
          int r1 = wrap(f1 - (int)(short)(off & 0xffff), d), r2 = wrap(f2 - (off >> 16), d);
          int4 h = hq[(r2 * d + r1) * HP4 + part];
          int4 w = Wt[q], w2 = s_w2[part];
          acc += (long long)w2.x * (long long)(max(h.x + sgn * w.x, 0) - max(h.x, 0));
          acc += (long long)w2.y * (long long)(max(h.y + sgn * w.y, 0) - max(h.y, 0));
          acc += (long long)w2.z * (long long)(max(h.z + sgn * w.z, 0) - max(h.z, 0));
          acc += (long long)w2.w * (long long)(max(h.w + sgn * w.w, 0) - max(h.w, 0));
        }
      }
      acc = warp_sum_ll(acc);
      const double delta = so_real(sv.c, sv.b2, sv.y_scale, sv.y_mean, acc >> SO_SPLIT_BITS, acc & SO_SPLIT_MASK, 0);
      bool guard;
      bool accept = metropolis(delta, T, u, P.exact_guard ? P.guard_tol : 0.0, guard);
      if (guard) {
        double de = so_exact_delta(sv, bits, d, f, lane);
        bool g2;
        accept = metropolis(de, T, u, 0.0, g2);
        ++n_guard;
      }
      if (accept) {
        if (P.type_hist) so_track_flip(P.type_hist, P.n_chains, c, bits, d, P.N, f, lane);        // before the bit flips
        if (MAXIT > 0) {
#pragma unroll
          for (int it = 0; it < MAXIT; ++it)
            if (lane + 32 * it < Q) hq[idx[it]] = hn[it];
        } else {
          for (int q = lane; q < Q; q += 32) {
            int k = q / HP4, part = q - k * HP4;
            int off = offt[k];
            int r1 = wrap(f1 - (int)(short)(off & 0xffff), d), r2 = wrap(f2 - (off >> 16), d);
            int4 *p = hq + (r2 * d + r1) * HP4 + part;
            int4 h = *p, w = Wt[q];
            h.x += sgn * w.x; h.y += sgn * w.y; h.z += sgn * w.z; h.w += sgn * w.w;
            *p = h;
          }
        }
        if (lane == 0) bits[f >> 5] = word ^ (1u << (f & 31));
        lo += acc & SO_SPLIT_MASK;
        hi += (acc >> SO_SPLIT_BITS) + (lo >> SO_SPLIT_BITS);
        lo &= SO_SPLIT_MASK;
        ++n_acc;
      }
      if (P.accept_out && lane == 0) P.accept_out[c * P.n_steps + s] = accept ? 1 : 0;
      __syncwarp();
    }
    if (lane == 0) {
      P.hi[c] = hi;
      P.lo[c] = lo;
      if (n_acc) atomicAdd(P.counters + 1, (unsigned long long)n_acc);
      if (n_guard) atomicAdd(P.counters + 2, (unsigned long long)n_guard);
    }
  }
}

// ---- gated kernel: tree gate on.  One warp per chain, one affected row per lane-slot ---------------------
template <int HP4>
__global__ void __launch_bounds__(128) sa_gated_kernel(SaParams P) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const SurView &sv = P.sv;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t *s_bits = (uint32_t *)s_raw + warp * P.W;
  const int d = P.d;
  int4 w2[HP4];
#pragma unroll
  for (int p = 0; p < HP4; ++p) w2[p] = __ldg((const int4 *)sv.w2q + p);

  for (;;) {
    long long c = 0;
    if (lane == 0) c = (long long)atomicAdd(P.counters, 1ULL);
    c = __shfl_sync(SO_FULL, c, 0);
    if (c >= P.n_chains) break;
    uint32_t *bits = P.bits + c * P.W;
    for (int w = lane; w < P.W; w += 32) s_bits[w] = bits[w];
    __syncwarp();
    int4 *hq = (int4 *)P.hq + c * (long long)P.N * HP4;
    long long hi = P.hi[c], lo = P.lo[c];
    int nact = P.nact[c];
    const double tsc = P.tscale[c];
    unsigned n_acc = 0, n_guard = 0;

    for (long long s = 0; s < P.n_steps; ++s) {
      const double T = __dmul_rn(tsc, __ldg(P.temps + s));
      int f;
      float u;
      draw(P, c, s, f, u);
      const int f2 = f / d, f1 = f - f2 * d;
      const int sgn = ((s_bits[f >> 5] >> (f & 31)) & 1) ? -1 : 1;
      BitView b0{s_bits, d, -1}, b1{s_bits, d, f};
      long long dhi = 0, dlo = 0;
      int dn = 0;
      for (int k = lane; k < sv.K; k += 32) {
        int off = __ldg(sv.offs + k);
        int r1 = wrap(f1 - (int)(short)(off & 0xffff), d), r2 = wrap(f2 - (off >> 16), d);
        const int4 *hrow = hq + (r2 * d + r1) * HP4;
        long long A0 = 0, A1 = 0;
#pragma unroll
        for (int p = 0; p < HP4; ++p) {
          int4 h = hrow[p];
          int4 w = __ldg((const int4 *)sv.W1q + k * HP4 + p);
          A0 += (long long)w2[p].x * max(h.x, 0) + (long long)w2[p].y * max(h.y, 0) +
                (long long)w2[p].z * max(h.z, 0) + (long long)w2[p].w * max(h.w, 0);
          A1 += (long long)w2[p].x * max(h.x + sgn * w.x, 0) + (long long)w2[p].y * max(h.y + sgn * w.y, 0) +
                (long long)w2[p].z * max(h.z + sgn * w.z, 0) + (long long)w2[p].w * max(h.w + sgn * w.w, 0);
        }
        int g0 = 1, g1 = 1;
        if (sv.gated) tree_gate_pair(sv, s_bits, d, r1, r2, f, g0, g1);
        if (g1) { dhi += A1 >> SO_SPLIT_BITS; dlo += A1 & SO_SPLIT_MASK; }
        if (g0) { dhi -= A0 >> SO_SPLIT_BITS; dlo -= A0 & SO_SPLIT_MASK; }
        dn += g1 - g0;
      }
      dhi = warp_sum_ll(dhi);
      dlo = warp_sum_ll(dlo);
      dn = warp_sum_i(dn);
      const double delta = so_real(sv.c, sv.b2, sv.y_scale, sv.y_mean, dhi, dlo, dn);
      bool guard;
      bool accept = metropolis(delta, T, u, P.exact_guard ? P.guard_tol : 0.0, guard);
      if (guard) {
        double de = so_exact_delta(sv, s_bits, d, f, lane);
        bool g2;
        accept = metropolis(de, T, u, 0.0, g2);
        ++n_guard;
      }
      if (accept) {
        if (P.type_hist) so_track_flip(P.type_hist, P.n_chains, c, s_bits, d, P.N, f, lane);      // before the bit flips
        for (int k = lane; k < sv.K; k += 32) {
          int off = __ldg(sv.offs + k);
          int r1 = wrap(f1 - (int)(short)(off & 0xffff), d), r2 = wrap(f2 - (off >> 16), d);
          int4 *hrow = hq + (r2 * d + r1) * HP4;
#pragma unroll
          for (int p = 0; p < HP4; ++p) {
            int4 h = hrow[p];
            int4 w = __ldg((const int4 *)sv.W1q + k * HP4 + p);
            h.x += sgn * w.x; h.y += sgn * w.y; h.z += sgn * w.z; h.w += sgn * w.w;
            hrow[p] = h;
          }
        }
        __syncwarp();
        if (lane == 0) s_bits[f >> 5] ^= (1u << (f & 31));
        lo += dlo;
        hi += dhi + (lo >> SO_SPLIT_BITS);
        lo &= SO_SPLIT_MASK;
        nact += dn;
        ++n_acc;
      }
      if (P.accept_out && lane == 0) P.accept_out[c * P.n_steps + s] = accept ? 1 : 0;
      __syncwarp();
    }
    for (int w = lane; w < P.W; w += 32) bits[w] = s_bits[w];
    if (lane == 0) {
      P.hi[c] = hi;
      P.lo[c] = lo;
      P.nact[c] = nact;
      if (n_acc) atomicAdd(P.counters + 1, (unsigned long long)n_acc);
      if (n_guard) atomicAdd(P.counters + 2, (unsigned long long)n_guard);
    }
    __syncwarp();
  }
}

// ---- apply explicit flips (flip_atom): bits only; the caller re-scores the touched chains -------------------
__global__ void apply_flips_kernel(long long n, const long long *chains, const int32_t *sites, uint32_t *bits_all,
                                   int W) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (long long i = 0; i < n; ++i) {   // sequential: repeated (chain, site) pairs toggle repeatedly
    int f = sites[i];
    bits_all[chains[i] * W + (f >> 5)] ^= (1u << (f & 31));
  }
}

// ---- best structure: arg-max of the tracked objective (ties -> lowest chain) ----------------------------------
__global__ void __launch_bounds__(1024) best_kernel(SurView sv, const long long *hi, const long long *lo,
                                                    const int32_t *nact, const uint32_t *bits, long long n, int W,
                                                    long long chain_id0, unsigned char *record) {
  __shared__ double s_of[32];
  __shared__ long long s_ix[32];
  __shared__ long long s_best;
  double best = -INFINITY;
  long long bi = -1;
  for (long long t = threadIdx.x; t < n; t += blockDim.x) {
    double v = so_real(sv.c, sv.b2, sv.y_scale, sv.y_mean, hi[t], lo[t], nact[t]);
    if (bi < 0 || v > best) { best = v; bi = t; }
  }
  if (bi < 0) bi = 0x7fffffffffffffffLL;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(SO_FULL, best, o);
    long long oi = __shfl_xor_sync(SO_FULL, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_of[warp] = best; s_ix[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    int nw = blockDim.x >> 5;
    best = lane < nw ? s_of[lane] : -INFINITY;
    bi = lane < nw ? s_ix[lane] : 0x7fffffffffffffffLL;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double ov = __shfl_xor_sync(SO_FULL, best, o);
      long long oi = __shfl_xor_sync(SO_FULL, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) {
      s_best = bi;
      *(double *)record = best;
      *(long long *)(record + 8) = chain_id0 + bi;
    }
  }
  __syncthreads();
  uint32_t *out = (uint32_t *)(record + 16);
  long long b = s_best;
  for (int w = threadIdx.x; w < W; w += blockDim.x) out[w] = bits[b * W + w];
}

// ---- replica exchange: one thread per temperature ladder (<= 32 rungs) -------------------------------------
// Global chain g = i*world + rank lives at position (g % world)*n_local + g/world of the all-gathered arrays.
// Chains of a ladder are ranked by their current temperature (hottest first, ties by global id); adjacent
// rungs (r, r+1), r = parity (mod 2), swap TEMPERATURES with probability
// min(1, exp[(OF_a - OF_b) * (1/T_b - 1/T_a)]) (maximisation, pi ~ exp(OF/T)).
__global__ void exchange_kernel(long long n_local, int world, int ladder, int parity, unsigned long long seed,
                                long long round, double t_ref, const double *of_all, const double *tscale_all,
                                double *tscale_new) {
  long long lad = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long n_global = n_local * world;
  if (lad >= n_global / ladder) return;
  long long base = lad * ladder;
  long long pos[32];
  int order[32];
  for (int i = 0; i < ladder; ++i) {
    long long g = base + i;
    pos[i] = (g % world) * n_local + g / world;
  }
  for (int i = 0; i < ladder; ++i) {   // insertion sort, descending temperature, stable
    int j = i;
    double ti = tscale_all[pos[i]];
    while (j > 0 && tscale_all[pos[order[j - 1]]] < ti) { order[j] = order[j - 1]; --j; }
    order[j] = i;
  }
  for (int r = parity; r + 1 < ladder; r += 2) {
    long long a = pos[order[r]], b = pos[order[r + 1]];
    double Ta = tscale_all[a] * t_ref, Tb = tscale_all[b] * t_ref;
    double arg = (of_all[a] - of_all[b]) * (1.0 / Tb - 1.0 / Ta);
    uint32_t o[4];
    philox4x32_10((uint32_t)lad, (uint32_t)((unsigned long long)lad >> 32), (uint32_t)round, (uint32_t)r,
                  (uint32_t)seed ^ 0x5bd1e995u, (uint32_t)(seed >> 32), o);
    double u = (double)(o[0] >> 8) * 5.9604644775390625e-08;
    bool swap = (arg >= 0.0) || (exp(arg) > u);
    if (swap) {
      tscale_new[a] = tscale_all[b];
      tscale_new[b] = tscale_all[a];
    }
  }
}

__global__ void exchange_commit_kernel(long long n_global, long long pos0, long long n_local, double *tscale_all,
                                       const double *tscale_new, double *tscale_local, unsigned long long *n_swapped) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n_global) return;
  double v = tscale_new[t];
  bool changed = v != tscale_all[t];
  tscale_all[t] = v;
  if (t >= pos0 && t < pos0 + n_local) {
    tscale_local[t - pos0] = v;
    if (changed) atomicAdd(n_swapped, 1ULL);
  }
}

// ---- launch wrappers -------------------------------------------------------------------------------------
#define SO_DISPATCH_HP4(hp4, ...)                                          \
  switch (hp4) {                                                           \
    case 1: { constexpr int HP4_ = 1; __VA_ARGS__; } break;                \
    case 2: { constexpr int HP4_ = 2; __VA_ARGS__; } break;                \
    case 3: { constexpr int HP4_ = 3; __VA_ARGS__; } break;                \
    case 4: { constexpr int HP4_ = 4; __VA_ARGS__; } break;                \
    case 5: { constexpr int HP4_ = 5; __VA_ARGS__; } break;                \
    case 6: { constexpr int HP4_ = 6; __VA_ARGS__; } break;                \
    case 7: { constexpr int HP4_ = 7; __VA_ARGS__; } break;                \
    case 8: { constexpr int HP4_ = 8; __VA_ARGS__; } break;                \
    default: so_set_error("hidden width %d not supported (max 32)", (hp4) * 4); return SO_EINVAL; \
  }

static const size_t kSmemTabLimit = 96 * 1024;

int so_launch_score(const so_lattice *lat, const so_surrogate *s, const uint32_t *bits_dev, int64_t n, int32_t *hq_dev,
                    long long *hi, long long *lo, int32_t *nact, cudaStream_t st) {
  if (n == 0) return SO_OK;
  {
    // rectangular-window surrogates: first layer on the tensor cores (so_score_tc.cu); SO_SCORE_TC=0 keeps the
    // CUDA-core kernel (both give the same bits; tests compare them)
    const char *env = getenv("SO_SCORE_TC");
    if (s->tc_ok && !(env && env[0] == '0')) return so_launch_score_tc(lat, s, bits_dev, n, hq_dev, hi, lo, nact, st);
  }
  SO_CUDA(cudaMemsetAsync(hi, 0, n * sizeof(long long), st));
  SO_CUDA(cudaMemsetAsync(lo, 0, n * sizeof(long long), st));
  SO_CUDA(cudaMemsetAsync(nact, 0, n * sizeof(int32_t), st));
  SurView sv = make_view(s);
  size_t tab = (size_t)s->K * s->Hp * 4 + (size_t)s->K * 4;
  int tab_in_smem = tab <= kSmemTabLimit;
  size_t smem = ((lat->W * 4 + 15) / 16) * 16 + (tab_in_smem ? tab : 0);
  int threads = lat->N >= 256 ? 256 : ((lat->N + 31) / 32) * 32;
  dim3 grid((lat->N + threads - 1) / threads, 1, 1);
  // gridDim.y is limited to 65535: loop over slabs of chains
  for (int64_t c0 = 0; c0 < n; c0 += 32768) {
    int64_t nc = n - c0 < 32768 ? n - c0 : 32768;
    grid.y = (unsigned)nc;
    SO_DISPATCH_HP4(s->HP4, {
      auto kern = score_kernel<HP4_>;
      SO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<grid, threads, smem, st>>>(bits_dev + c0 * lat->W, lat->d, lat->N, lat->W, sv, tab_in_smem,
                                        hq_dev ? hq_dev + c0 * (long long)lat->N * s->Hp : nullptr, hi + c0, lo + c0,
                                        nact + c0);
    });
    SO_CUDA(cudaGetLastError());
  }
  normalize_sums_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(hi, lo, n);
  SO_CUDA(cudaGetLastError());
  return SO_OK;
}

int so_normalize_sums(long long *hi, long long *lo, int64_t n, cudaStream_t st) {
  if (n == 0) return SO_OK;
  normalize_sums_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(hi, lo, n);
  SO_CUDA(cudaGetLastError());
  return SO_OK;
}

int so_launch_objective(const so_surrogate *s, const long long *hi, const long long *lo, const int32_t *nact, int64_t n,
                        double *of_dev, cudaStream_t st) {
  if (n == 0) return SO_OK;
  objective_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(make_view(s), hi, lo, nact, n, of_dev);
  SO_CUDA(cudaGetLastError());
  return SO_OK;
}

int so_launch_eval_rows(const so_surrogate *s, int64_t n_rows, const uint8_t *X_dev, uint8_t *active_dev,
                        double *rate_dev, cudaStream_t st) {
  if (n_rows == 0) return SO_OK;
  SurView sv = make_view(s);
  SO_DISPATCH_HP4(s->HP4, {
    eval_rows_kernel<HP4_><<<(unsigned)((n_rows + 127) / 128), 128, 0, st>>>(sv, n_rows, X_dev, active_dev, rate_dev);
  });
  SO_CUDA(cudaGetLastError());
  return SO_OK;
}

template <int HP4, int MAXIT, bool TAB_SMEM>
static int launch_flat(const SaParams &P, size_t smem, int blocks, cudaStream_t st) {
  auto kern = sa_flat_kernel<HP4, MAXIT, TAB_SMEM>;
  SO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<blocks, 256, smem, st>>>(P);
  SO_CUDA(cudaGetLastError());
  return SO_OK;
}

int so_launch_anneal(const so_chains *c, const SaParams &P, cudaStream_t st, int *kernel_id) {
  int kid_local = 0;
  int &kid = kernel_id ? *kernel_id : kid_local;
  const so_surrogate *s = c->sur;
  int dev = c->lat->device, n_sm = 0;
  SO_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  SO_CUDA(cudaMemsetAsync(P.counters, 0, sizeof(unsigned long long), st));   // work counter
  // row records of the gated window kernel: current only if that kernel ran last and nothing else touched the bits since
  const int records_current = c->meta_valid;
  c->meta_valid = 0;
  if (s->compact) {          // 24-bit packed state: the pipelined window kernel is the only reader
    if (!(P.variant == 0 || P.variant == 3)) {
      so_set_error("the compact state is served by the window kernel only (variant 0 or 3, got %d)", P.variant);
      return SO_EINVAL;
    }
    if (s->gated) {
      int rc = so_launch_anneal_window_gated(c, P, st, !records_current);
      if (rc == SO_OK) { kid = SO_K_WINDOW_GATED; c->meta_valid = 1; }
      return rc;
    }
    return kid = SO_K_WINDOW, so_launch_anneal_window(c, P, st);
  }
  // small lattices: the whole chain state stays in shared memory for the launch (so_sa_resident.cu)
  // few chains on a small lattice: a CTA per chain (latency); many: a warp per chain (throughput)
  if ((P.variant == 0 || P.variant == 5) && so_team_ok(s, P.N, P.W, P.variant == 5 ? 1 : P.n_chains, n_sm))
    return kid = SO_K_TEAM, so_launch_anneal_team(c, P, st);
  if ((P.variant == 0 && so_resident_plan(s, P.N, P.W, nullptr, nullptr)) ||
      (P.variant == 4 && so_resident_plan(s, P.N, P.W, nullptr, nullptr, 1)))
    return kid = SO_K_RESIDENT, so_launch_anneal_resident(c, P, st);
  // gated window descriptor on a large lattice: gate records ride the TMA pipeline of the window kernel (so_sa_window_gated.cu)
  if ((P.variant == 0 || P.variant == 3) && so_window_gated_ok(c)) {
    int rc = so_launch_anneal_window_gated(c, P, st, !records_current);
    if (rc == SO_OK) { kid = SO_K_WINDOW_GATED; c->meta_valid = 1; }
    return rc;
  }
  // gated, any descriptor: cached gates + path index, only the active rows are read (so_sa_resident.cu); variant 6 asks for it
  if ((P.variant == 0 || P.variant == 3 || P.variant == 6) && so_gate_index_ok(s, P.N, P.W))
    return kid = SO_K_GATE_INDEX, so_launch_anneal_gate_index(c, P, st);
  if (!s->gated && (P.variant == 0 || P.variant == 3) && s->win_ok && s->K * s->HP4 <= 256 &&
      so_window_smem_bytes(s->K * s->HP4, P.W, 3) <= 113 * 1024) {
    return kid = SO_K_WINDOW, so_launch_anneal_window(c, P, st);
  }
  if (!s->gated && P.variant != 1) {
    kid = SO_K_FLAT;
    int Q = s->K * s->HP4;
    size_t tab = (size_t)Q * 16 + (size_t)s->K * 4;
    bool tab_smem = tab <= kSmemTabLimit;
    size_t smem = (tab_smem ? tab : 0) + (size_t)s->HP4 * 16 + 16;
    int maxit = (Q + 31) / 32;
    // persistent grid: enough warps to cover the HBM latency-bandwidth product, chains handed out dynamically
    long long want = (P.n_chains + 7) / 8;
    int blocks = (int)(want < (long long)n_sm * 6 ? want : (long long)n_sm * 6);
    if (blocks < 1) blocks = 1;
    SO_DISPATCH_HP4(s->HP4, {
      if (tab_smem && maxit <= 8) return launch_flat<HP4_, 8, true>(P, smem, blocks, st);
      if (tab_smem) return launch_flat<HP4_, 0, true>(P, smem, blocks, st);
      return launch_flat<HP4_, 0, false>(P, smem, blocks, st);
    });
  } else {
    kid = SO_K_GATED;
    size_t smem = (size_t)4 * P.W * 4;
    long long want = (P.n_chains + 3) / 4;
    int blocks = (int)(want < (long long)n_sm * 8 ? want : (long long)n_sm * 8);
    if (blocks < 1) blocks = 1;
    SO_DISPATCH_HP4(s->HP4, {
      auto kern = sa_gated_kernel<HP4_>;
      SO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<blocks, 128, smem, st>>>(P);
    });
    SO_CUDA(cudaGetLastError());
  }
  return SO_OK;
}

int so_launch_apply_flips(const so_chains *c, int64_t n, const int64_t *chains_dev, const int32_t *sites_dev,
                          cudaStream_t st) {
  if (n == 0) return SO_OK;
  apply_flips_kernel<<<1, 32, 0, st>>>(n, (const long long *)chains_dev, sites_dev, c->bits_dev, c->lat->W);
  SO_CUDA(cudaGetLastError());
  return SO_OK;
}

int so_launch_best(const so_chains *c, int64_t chain_id0, void *record_dev, cudaStream_t st) {
  best_kernel<<<1, 1024, 0, st>>>(make_view(c->sur), c->hi_dev, c->lo_dev, c->nact_dev, c->bits_dev, c->n_chains,
                                  c->lat->W, chain_id0, (unsigned char *)record_dev);
  SO_CUDA(cudaGetLastError());
  return SO_OK;
}

int so_launch_exchange(const so_chains *c, int rank, int world, int ladder, int parity, uint64_t seed, int64_t round,
                       double t_ref, const double *of_all, double *tscale_all, double *scratch,
                       unsigned long long *n_swapped, cudaStream_t st) {
  long long n_global = c->n_chains * (long long)world;
  SO_CUDA(cudaMemcpyAsync(scratch, tscale_all, n_global * sizeof(double), cudaMemcpyDeviceToDevice, st));
  long long ladders = n_global / ladder;
  if (ladders > 0) {
    exchange_kernel<<<(unsigned)((ladders + 127) / 128), 128, 0, st>>>(c->n_chains, world, ladder, parity, seed, round,
                                                                       t_ref, of_all, tscale_all, scratch);
    SO_CUDA(cudaGetLastError());
  }
  exchange_commit_kernel<<<(unsigned)((n_global + 255) / 256), 256, 0, st>>>(
      n_global, (long long)rank * c->n_chains, c->n_chains, tscale_all, scratch, c->tscale_dev, n_swapped);
  SO_CUDA(cudaGetLastError());
  return SO_OK;
}
