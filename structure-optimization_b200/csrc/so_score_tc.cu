// so_score_tc.cu -- first-layer state + objective of whole structures on the 5th-generation tensor cores (tcgen05).
//
// Same result as score_kernel (so_sa.cu), bit for bit: the first layer  hq[r][j] = b1q[j] + sum_k x[r+off_k] W1q[k][j]
// is an integer GEMM  X[rows x K] (0/1 bytes) * W[K x 4*Hp]  where every int32 weight is split into four BALANCED 8-bit
// limbs (w = l0 + 2^8 l1 + 2^16 l2 + 2^24 l3, l in [-128,127]); kind::i8 accumulates the limb products exactly in
// int32 in TMEM and the epilogue recombines them with shifts.  ncu of score_kernel shows it bound by integer issue
// (980 IMAD per row), which is what makes the contraction pay.
//
//   * one CTA (128 threads) walks the 128-row tiles of one chain; a thread owns one lattice site of the tile;
//   * the descriptor row of a site (rectangular window, <= 8 x 8) is gathered from the chain's bit-plane with one
//     funnel shift per window row and expanded to bytes through a 256-entry look-up table -> A tile in shared memory
//     in the canonical K-major no-swizzle layout (8 x 16 B core matrices), 64 K-bytes per row;
//   * one thread issues 2 x tcgen05.mma (M=128, N=4*Hp, K=32) and commits to an mbarrier; the four warps read their
//     TMEM lane quarter back with tcgen05.ld, recombine limbs, apply relu . w2q (exact int64) and store the hq row.
#include "so_internal.cuh"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor): start>>4 [0,14),
// leading byte offset>>4 [16,30) (between the two 16-B K chunks), stride byte offset>>4 [32,46) (between 8-row
// groups), version 1 at [46,48), layout type 0 at [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);
}

__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

struct TcGeom {
  int a1, a2;        // smallest offsets of the window; K slot p2*KW + p1 <-> site (r1 + a1 + p1, r2 + a2 + p2)
  int L, n_seg;      // window extent (<= 16 each)
  int kw;            // K bytes per window row: 8 (L <= 8) or 16
  int n_chunks;      // 16-B K chunks per row (even: one MMA consumes two)
  int n_mma;         // 4 * Hp, a multiple of 16
  int b_lbo;         // byte distance between 16-B K chunks of the B image = n_mma * 16
};

constexpr int TILE_M = 128;
constexpr int A_LBO = TILE_M * 16;          // 2048 B between 16-B K chunks of the A tile

// L (<= 16) consecutive bits of a bit-plane starting at site (i1s .. i1s+L-1, i2), periodic in i1
__device__ __forceinline__ uint32_t window_bits(const uint32_t *bits, int d, int i1s, int i2, int L) {
  auto run = [&](int v0, int n) -> uint32_t {          // n <= 16 bits starting at site index v0
    uint32_t w0 = bits[v0 >> 5], w1 = bits[(v0 >> 5) + 1];   // (the bit-plane is padded by one word)
    return __funnelshift_r(w0, w1, v0 & 31) & ((1u << n) - 1u);
  };
  if (i1s + L <= d) return run(i2 * d + i1s, L);
  int n1 = d - i1s;
  return run(i2 * d + i1s, n1) | (run(i2 * d, L - n1) << n1);
}

// Rows are the concatenation (chain, site): global row G = chain * N + site; a CTA owns a contiguous range of
// 128-row tiles, so a tile touches at most two chains (N >= 128) whose bit-planes sit in two shared-memory slots.
// NCH = 16-B K chunks per row at compile time (4 = the <= 8 x 8 window case, loops fully unrolled) or 0 = runtime.
template <int HP4, int NCH>
__global__ void __launch_bounds__(128) score_tc_kernel(const uint32_t *__restrict__ bits_all, int d, int N, int W,
                                                       long long n_chains, SurView sv, TcGeom g,
                                                       const uint4 *__restrict__ b_image, int32_t *hq_all,
                                                       long long *hi_out, long long *lo_out, int32_t *nact_out,
                                                       const int32_t *__restrict__ bias24) {
  // bias24 != null (HP4 = 5 only): hq_all is the 24-bit packed state, 64 bytes per row (so_sa_window.cu: pack24_kernel's layout);
  // the epilogue packs the row itself and the int32 slab + pack pass of the state build disappear
  constexpr int HP = HP4 * 4;
  const int n_chunks = NCH ? NCH : g.n_chunks;
  const bool kw8 = NCH ? true : (g.kw == 8);
  extern __shared__ __align__(128) unsigned char s_raw[];
  unsigned char *s_A = s_raw;                                   // [n_chunks][16 groups][8 rows][16 B]
  unsigned char *s_B = s_A + n_chunks * A_LBO;                  // [n_chunks][n_mma/8 groups][8][16 B]
  uint2 *s_lut = (uint2 *)(s_B + n_chunks * g.b_lbo);           // 256 x 8 B: bit pattern -> bytes
  int4 *s_out = (int4 *)(s_lut + 256);                          // 2 x [128 rows][HP4 int4]: hq tile staged for a bulk store
  uint32_t *s_bits = (uint32_t *)(s_out + 2 * TILE_M * HP4);    // 2 slots x (W + 1) words
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < (n_chunks * g.b_lbo) / 16; i += blockDim.x) ((uint4 *)s_B)[i] = b_image[i];
  for (int i = tid; i < 256; i += blockDim.x) {
    uint32_t lo = 0, hi = 0;
    for (int b = 0; b < 4; ++b) { lo |= ((i >> b) & 1u) << (8 * b); hi |= ((i >> (b + 4)) & 1u) << (8 * b); }
    s_lut[i] = make_uint2(lo, hi);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {                                             // 128 accumulator columns (n_mma <= 128 here)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  // instruction descriptor (cute::UMMA::InstrDescriptor): D = s32 [4,6)=2, A = u8 [7,10)=0, B = s8 [10,13)=1,
  // both K-major, N>>3 at [17,23), M>>4 at [24,29)
  const uint32_t idesc = (2u << 4) | (1u << 10) | ((uint32_t)(g.n_mma >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
  const uint32_t a_addr = smem_u32(s_A), b_addr = smem_u32(s_B), bar = smem_u32(&s_bar);
  uint32_t parity = 0;
  int b1q[HP], w2q[HP];
#pragma unroll
  for (int j = 0; j < HP; ++j) { b1q[j] = __ldg(sv.b1q + j); w2q[j] = __ldg(sv.w2q + j); }

  const long long total_rows = n_chains * (long long)N;
  const long long n_tiles = (total_rows + TILE_M - 1) / TILE_M;
  const long long per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
  const long long t_begin = blockIdx.x * per_cta, t_end = min(n_tiles, t_begin + per_cta);
  long long slot_chain = -1;                        // chain whose bit-plane is in slot 0 (slot 1 = the next chain)
  int buf = 0;
  long long c0 = (t_begin * TILE_M) / N;            // chain and site of the first row of the current tile,
  int r0 = (int)(t_begin * TILE_M - c0 * N);        // advanced incrementally (N >= 128: at most one wrap per tile)

  for (long long t = t_begin; t < t_end; ++t) {
    const long long G0 = t * TILE_M, G = G0 + tid;
    if (c0 != slot_chain) {                         // (re)load the two bit-planes this and the following tiles need
      __syncthreads();
      for (int w = tid; w < 2 * (W + 1); w += blockDim.x) {
        int sl = w / (W + 1), ww = w - sl * (W + 1);
        long long cc = c0 + sl;
        s_bits[w] = (ww < W && cc < n_chains) ? bits_all[cc * W + ww] : 0u;
      }
      slot_chain = c0;
      __syncthreads();
    }
    const bool live = G < total_rows;
    const int rr = r0 + tid;
    const int sl_mine = (live && rr >= N) ? 1 : 0;  // 0: chain c0, 1: chain c0 + 1
    const int r = live ? (sl_mine ? rr - N : rr) : 0;
    const int r2 = r / d, r1 = r - r2 * d;
    const uint32_t *mybits = s_bits + (size_t)sl_mine * (W + 1);
    // ---- A tile: this thread's descriptor row, kw K-bytes per window row -------------------------------------------
    {
      unsigned char *row = s_A + (tid >> 3) * 128 + (tid & 7) * 16;
      const int i1s = wrap(r1 + g.a1, d);
      if (kw8) {
#pragma unroll
        for (int ck = 0; ck < (NCH ? NCH : 16); ++ck) {
          if (ck >= n_chunks) break;
          uint2 s0 = make_uint2(0u, 0u), s1 = make_uint2(0u, 0u);
          if (live && 2 * ck < g.n_seg) s0 = s_lut[window_bits(mybits, d, i1s, wrap(r2 + g.a2 + 2 * ck, d), g.L)];
          if (live && 2 * ck + 1 < g.n_seg) s1 = s_lut[window_bits(mybits, d, i1s, wrap(r2 + g.a2 + 2 * ck + 1, d), g.L)];
          *(uint4 *)(row + ck * A_LBO) = make_uint4(s0.x, s0.y, s1.x, s1.y);
        }
      } else {
        for (int ck = 0; ck < g.n_chunks; ++ck) {
          uint2 s0 = make_uint2(0u, 0u), s1 = make_uint2(0u, 0u);
          if (live && ck < g.n_seg) {
            uint32_t pat = window_bits(mybits, d, i1s, wrap(r2 + g.a2 + ck, d), g.L);
            s0 = s_lut[pat & 255u];
            s1 = s_lut[pat >> 8];
          }
          *(uint4 *)(row + ck * A_LBO) = make_uint4(s0.x, s0.y, s1.x, s1.y);
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> tensor-core reads
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int kk = 0; kk < (NCH ? NCH : 16); kk += 2) {
        if (kk >= n_chunks) break;
        mma_i8(tmem, umma_desc(a_addr + kk * A_LBO, A_LBO, 128), umma_desc(b_addr + kk * g.b_lbo, g.b_lbo, 128), idesc,
               kk > 0 ? 1u : 0u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    }
    {
      uint32_t ok;
      do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
      } while (!ok);
      parity ^= 1u;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- epilogue: TMEM lane = tile row; columns n = 4*j + limb ---------------------------------------------------
    long long my_hi = 0, my_lo = 0;
    int my_n = 0;
    {
      int h[HP];
      const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
      for (int q = 0; q < HP4; ++q) {
        int v[16];
        tmem_ld16(lane_base + q * 16, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int j = 4 * q + jj;
          h[j] = b1q[j] + v[4 * jj] + v[4 * jj + 1] * 256 + v[4 * jj + 2] * 65536 + (int)((unsigned)v[4 * jj + 3] << 24);
        }
      }
      if (live) {
        long long A = 0;
#pragma unroll
        for (int j = 0; j < HP; ++j) A += (long long)w2q[j] * (long long)max(h[j], 0);
        if (hq_all) {
          if (HP4 == 5 && bias24) {                   // five biased 24-bit values per int4, four int4 per row
            int4 *dst = s_out + (size_t)buf * TILE_M * HP4 + (size_t)tid * 4;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              const unsigned v0 = (unsigned)(h[5 * p] + __ldg(bias24 + 5 * p)) & 0xffffffu,
                             v1 = (unsigned)(h[5 * p + 1] + __ldg(bias24 + 5 * p + 1)) & 0xffffffu,
                             v2 = (unsigned)(h[5 * p + 2] + __ldg(bias24 + 5 * p + 2)) & 0xffffffu,
                             v3 = (unsigned)(h[5 * p + 3] + __ldg(bias24 + 5 * p + 3)) & 0xffffffu,
                             v4 = (unsigned)(h[5 * p + 4] + __ldg(bias24 + 5 * p + 4)) & 0xffffffu;
              dst[p] = make_int4((int)(v0 | (v1 << 24)), (int)((v1 >> 8) | (v2 << 16)), (int)((v2 >> 16) | (v3 << 8)), (int)v4);
            }
          } else {                                    // 80-B row stride: conflict-free 16-B shared stores
            int4 *dst = s_out + (size_t)buf * TILE_M * HP4 + (size_t)tid * HP4;
#pragma unroll
            for (int p = 0; p < HP4; ++p) dst[p] = make_int4(h[4 * p], h[4 * p + 1], h[4 * p + 2], h[4 * p + 3]);
          }
        }
        BitView bv{mybits, d, -1};
        const int gate = sv.gated ? tree_gate(sv, bv, r1, r2) : 1;
        if (gate) { my_hi = A >> SO_SPLIT_BITS; my_lo = A & SO_SPLIT_MASK; my_n = 1; }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // TMEM reads done before the next tile's MMA
    // objective sums: a tile touches chains c0 and c0 + 1; exact integer atomics, normalised by the caller
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
      const bool mine = live && sl_mine == sl;
      const long long vh = mine ? my_hi : 0;        // |hi| < 2^42: three 21-bit limbs; lo < 2^20, n <= 1: one REDUX each
      const long long shi = ((long long)__reduce_add_sync(SO_FULL, (int)(vh >> 42)) << 42) +
                            ((long long)__reduce_add_sync(SO_FULL, (int)((vh >> 21) & 0x1fffff)) << 21) +
                            (long long)__reduce_add_sync(SO_FULL, (int)(vh & 0x1fffff));
      const long long slo = (long long)__reduce_add_sync(SO_FULL, (int)(mine ? my_lo : 0));
      const int sn = __reduce_add_sync(SO_FULL, mine ? my_n : 0);
      if (lane == 0 && (shi | slo | sn) != 0 && c0 + sl < n_chains) {
        atomicAdd((unsigned long long *)(hi_out + c0 + sl), (unsigned long long)shi);
        atomicAdd((unsigned long long *)(lo_out + c0 + sl), (unsigned long long)slo);
        atomicAdd(nact_out + c0 + sl, sn);
      }
    }
    if (hq_all) {
      // the tile's rows are contiguous in hq: one bulk-async store of up to 128*4*HP bytes, double buffered
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        const int rows = (int)min((long long)TILE_M, total_rows - G0);
        const uint32_t row_bytes = (HP4 == 5 && bias24) ? 64u : (uint32_t)(HP * 4);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"((char *)hq_all + (size_t)G0 * row_bytes),
                     "r"(smem_u32(s_out + (size_t)buf * TILE_M * HP4)), "r"((uint32_t)rows * row_bytes)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the other buffer is free again
      }
      buf ^= 1;
    }
    r0 += TILE_M;
    if (r0 >= N) { r0 -= N; ++c0; }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

}  // namespace

// host: the B operand image (balanced 8-bit limbs of W1q in the canonical K-major layout) for a window surrogate.
// K slot s = p2*kw + p1 holds descriptor column k with offset (a1+p1, a2+p2); column n = 4*j + limb.
int so_build_tc_image(const so_surrogate *s, std::vector<unsigned char> &img, int *n_mma_out) {
  const int Hp = s->Hp, n_mma = 4 * Hp;
  const int L = s->win_b1 - s->win_a1 + 1, nseg = s->win_b2 - s->win_a2 + 1;
  if (!s->win_ok || L > 16 || nseg > 16 || n_mma > 128 || (n_mma % 16) != 0 || s->lat->N < TILE_M) return SO_EINVAL;
  const int kw = L <= 8 ? 8 : 16;
  const int n_chunks = ((nseg * kw + 31) / 32) * 2;
  const int b_lbo = n_mma * 16;
  img.assign((size_t)n_chunks * b_lbo, 0);
  for (int k = 0; k < s->K; ++k) {
    int o1 = (int)(short)(s->offs_host[k] & 0xffff), o2 = s->offs_host[k] >> 16;
    int slot = (o2 - s->win_a2) * kw + (o1 - s->win_a1);
    for (int j = 0; j < Hp; ++j) {
      long long w = s->W1q_host[(size_t)k * Hp + j];
      for (int l = 0; l < 4; ++l) {
        long long limb = l < 3 ? (((w + 128) & 255) - 128) : w;      // balanced digit; the top limb keeps the sign
        w = (w - limb) >> 8;
        if (l == 3 && (limb < -128 || limb > 127)) return SO_EINVAL;
        int n = 4 * j + l;
        img[(size_t)(slot >> 4) * b_lbo + (size_t)(n >> 3) * 128 + (size_t)(n & 7) * 16 + (slot & 15)] = (unsigned char)(signed char)limb;
      }
    }
  }
  *n_mma_out = n_mma;
  return SO_OK;
}

int so_launch_score_tc(const so_lattice *lat, const so_surrogate *s, const uint32_t *bits_dev, int64_t n, int32_t *hq_dev,
                       long long *hi, long long *lo, int32_t *nact, cudaStream_t st, int packed_out) {
  if (n == 0) return SO_OK;
  SO_REQUIRE(!packed_out || (s->compact && s->HP4 == 5 && s->bias24_dev), "packed output needs a compact surrogate");
  TcGeom g;
  g.a1 = s->win_a1; g.a2 = s->win_a2; g.L = s->win_b1 - s->win_a1 + 1; g.n_seg = s->win_b2 - s->win_a2 + 1;
  g.kw = g.L <= 8 ? 8 : 16;
  g.n_chunks = ((g.n_seg * g.kw + 31) / 32) * 2;
  g.n_mma = 4 * s->Hp; g.b_lbo = g.n_mma * 16;
  SO_CUDA(cudaMemsetAsync(hi, 0, n * sizeof(long long), st));
  SO_CUDA(cudaMemsetAsync(lo, 0, n * sizeof(long long), st));
  SO_CUDA(cudaMemsetAsync(nact, 0, n * sizeof(int32_t), st));
  size_t smem = (size_t)g.n_chunks * A_LBO + (size_t)g.n_chunks * g.b_lbo + 256 * 8 + (size_t)2 * TILE_M * s->Hp * 4 +
                (size_t)2 * (lat->W + 1) * 4 + 64;
  int n_sm = 0;
  SO_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, lat->device));
  long long tiles = (n * (long long)lat->N + TILE_M - 1) / TILE_M;
  int blocks = (int)(tiles < (long long)n_sm * 4 ? tiles : (long long)n_sm * 4);     // 4 CTAs x 128 TMEM columns per SM
  SurView sv = make_view(s);
  switch (s->HP4) {
#define SO_TC_CASE(V)                                                                                              \
  case V: {                                                                                                        \
    auto kern = (g.kw == 8 && g.n_chunks == 4) ? score_tc_kernel<V, 4> : score_tc_kernel<V, 0>;                    \
    SO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                   \
    kern<<<blocks, 128, smem, st>>>(bits_dev, lat->d, lat->N, lat->W, n, sv, g, (const uint4 *)s->tc_image_dev,    \
                                    hq_dev, hi, lo, nact, packed_out ? s->bias24_dev : nullptr);                   \
  } break;
    SO_TC_CASE(1) SO_TC_CASE(2) SO_TC_CASE(3) SO_TC_CASE(4) SO_TC_CASE(5) SO_TC_CASE(6) SO_TC_CASE(7) SO_TC_CASE(8)
#undef SO_TC_CASE
    default: so_set_error("hidden width not supported by the tensor-core scorer"); return SO_EINVAL;
  }
  SO_CUDA(cudaGetLastError());
  return so_normalize_sums(hi, lo, n, st);
}
