// so_sa_window_gated.cu -- the pipelined window kernel (so_sa_window.cu) WITH the classifier gate of the reference's surrogate
// (OML/train_surrogate.py:34-54: rows the decision tree calls inactive contribute nothing), for rectangular-window descriptors
// on lattices too large for the shared-memory-resident kernels.  Same arithmetic and decisions as every other SA kernel.
//
// What the gate needs per flip at site f, for each of the K rows r = f - off_k of the window:
//   g0(r) = gate of the row now, g1(r) = gate after the flip; rows with g0 = g1 = 1 enter the relu-difference sum, a row that
//   switches contributes its whole row sum (+ new / - old), the others nothing.
// The cached state that makes this cheap RIDES THE SAME TMA PIPELINE as the first-layer state: next to hq there is one 16-byte
// record per row, {T lo, T hi | g << 31, 0, 0}: g = the row's current gate, T = the set of window positions p at which the row
// sits when the flipped site is one its current tree path tests (bit p <=> the path tests column k(p)).  The window of a step is
// fetched as two tensor-TMA boxes on one mbarrier (state 64 or 80 bytes per site, records 16); lane p reads the record of the
// row at window position p: g0 is a bit, and the row needs a tree walk ONLY if bit p of its T is set (about depth / K of the
// rows).  Walks read the flipped structure from a (2L-1) x (2 n_seg - 1) bit patch around f built once per step in shared
// memory, so a node visit has no wrap arithmetic; nodes are 8 bytes (children for bit 0 / 1, patch-relative coordinates, the
// window position they mark).  On accept the walked rows' new records are written into the stage and go back with the same
// store as the state.  No path index, no gate bit-plane, no list upkeep.
// Records are (re)built by the kernel at the start of a launch unless the handle knows they are current (so_chains.meta_valid).
#include "so_window_common.cuh"
#include "so_types.cuh"
#include <stdlib.h>

namespace {

struct WinGate {
  int4 *meta;            // [n_chains][N] records
  const uint2 *tree;     // compact nodes (so_surrogate.wtree_dev)
  int n_tree;
  int tree_smem;         // 1: the nodes are copied to shared memory; 0: walked where they are (L1 / L2)
  int a1, a2;            // smallest offsets of the window
  int build;             // 1: build the records of every chain first
};

// node.x = child taken when the tested bit is 0 | child taken when it is 1 << 16
// node.y = u1 | u2 << 8 | pc << 16 | flags << 24: the tested column's offset minus the smallest offset (patch-relative), the
//          window position pc = (b2 - o2) * L + (b1 - o1) its rows sit at; flags: 1 leaf, 2 leaf is active, 4 the test depends on
//          its bit (thresholds outside [0, 1) send both values the same way and are not marked)
__device__ __forceinline__ unsigned shl_clamp(unsigned v, unsigned s) {   // v << s, 0 for s > 31 (PTX shl clamps)
  unsigned r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(s));
  return r;
}

// gate of the row at window position (p1, p2) on the patch (the structure AFTER the flip); one 8-byte node and one patch word
// per level, byte fields taken with PRMT
template <class TreePtr>
__device__ __forceinline__ unsigned walk_gate(TreePtr tree, const uint32_t *patch, int p1, int p2) {
  uint2 nd = tree[0];
  while (!(nd.y & 0x01000000u)) {
    const unsigned u1 = __byte_perm(nd.y, 0u, 0x4440), u2 = __byte_perm(nd.y, 0u, 0x4441);
    const unsigned bit = (patch[p2 + u2] >> (p1 + u1)) & 1u;
    nd = tree[(nd.x >> (bit << 4)) & 0xffffu];
  }
  return (nd.y >> 25) & 1u;
}
// the same walk, also collecting the window positions the path marks (accepted flips only): returns {T lo, T hi | g << 31}
__device__ __noinline__ uint2 walk_mark_cold(const uint2 *tree, const uint32_t *patch, int p1, int p2) {
  uint2 nd = tree[0];
  unsigned tlo = 0, thi = 0;
  while (!(nd.y & 0x01000000u)) {
    const unsigned u1 = nd.y & 0xffu, u2 = (nd.y >> 8) & 0xffu, pc = (nd.y >> 16) & 0xffu;
    const unsigned bit = (patch[p2 + u2] >> (p1 + u1)) & 1u;
    if (nd.y & 0x04000000u) { tlo |= shl_clamp(1u, pc); thi |= shl_clamp(1u, pc - 32u); }
    nd = tree[(nd.x >> (bit << 4)) & 0xffffu];
  }
  return make_uint2(tlo, thi | (((nd.y >> 25) & 1u) << 31));
}

// the same walk on the chain's bit-plane (record build at launch start): returns {T lo, T hi | g << 31}
__device__ __noinline__ uint2 walk_site_cold(const uint2 *tree, const uint32_t *bits, int d, int r1, int r2, int a1, int a2) {
  uint2 nd = tree[0];
  unsigned tlo = 0, thi = 0;
  while (!(nd.y & 0x01000000u)) {
    const int u1 = nd.y & 0xffu, u2 = (nd.y >> 8) & 0xffu;
    const unsigned pc = (nd.y >> 16) & 0xffu;
    const int v = wrap(r2 + u2 + a2, d) * d + wrap(r1 + u1 + a1, d);
    const unsigned bit = (bits[v >> 5] >> (v & 31)) & 1u;
    if (nd.y & 0x04000000u) { tlo |= shl_clamp(1u, pc); thi |= shl_clamp(1u, pc - 32u); }
    nd = tree[bit ? (nd.x >> 16) : (nd.x & 0xffffu)];
  }
  return make_uint2(tlo, thi | (((nd.y >> 25) & 1u) << 31));
}

// n <= 31 bits of the bit-plane starting at bit s (the word behind the plane exists: one pad word per warp)
__device__ __forceinline__ unsigned extract_bits(const uint32_t *bits, int s, int n) {
  return __funnelshift_r(bits[s >> 5], bits[(s >> 5) + 1], s & 31) & ((1u << n) - 1u);
}
// the common case, a patch no wider than the lattice: one or two pieces, and the flipped site appears exactly once (column L - 1
// of the centre row, toggled by the caller)
__device__ __forceinline__ unsigned patch_row_narrow(const uint32_t *bits, int d, int y, int x0, int PW) {
  const int n1 = min(PW, d - x0);
  unsigned r = extract_bits(bits, y * d + x0, n1);
  if (n1 < PW) r |= extract_bits(bits, y * d, PW - n1) << n1;
  return r;
}
// row `y` of the patch: PW sites from column x0 on, cyclic, with site (f1, f2) read inverted wherever it appears
__device__ __noinline__ unsigned patch_row(const uint32_t *bits, int d, int y, int x0, int PW, int f1, int f2) {
  unsigned r = 0;
  int j = 0, x = x0;
  while (j < PW) {
    const int n = min(PW - j, d - x);
    r |= extract_bits(bits, y * d + x, n) << j;
    j += n;
    x = 0;
  }
  if (y == f2) {
    int j0 = f1 - x0;
    if (j0 < 0) j0 += d;
    for (; j0 < PW; j0 += d) r ^= 1u << j0;
  }
  return r;
}

// fp64 side of the decision for a given delta; bit 0 accept, bit 1 guard
__device__ __noinline__ int decide_delta_cold(double delta, double tsc, const double *temps, long long s, float u,
                                              int exact_guard, double tol) {
  const double T = __dmul_rn(tsc, __ldg(temps + s));
  return metropolis_cold(delta, T, u, exact_guard, tol);
}

// whole row sum of one switching row, the part of this lane (lanes 0 .. SI4-1): sum_j w2_j relu(hq_j + w_j)
template <bool PACKED>
__device__ __forceinline__ long long row_part(const int4 h, const int4 w, int w5, const int4 w2, int w2_5, int B0, int B1,
                                              int B2, int B3, int B4) {
  if constexpr (PACKED) {
    const int v0 = h.x & 0xffffff;
    const int v1 = (int)(__funnelshift_r((unsigned)h.x, (unsigned)h.y, 24) & 0xffffffu);
    const int v2 = (int)(__funnelshift_r((unsigned)h.y, (unsigned)h.z, 16) & 0xffffffu);
    const int v3 = (int)((unsigned)h.z >> 8);
    const int v4 = h.w;
    return (long long)w2.x * (max(v0 + w.x, B0) - B0) + (long long)w2.y * (max(v1 + w.y, B1) - B1) +
           (long long)w2.z * (max(v2 + w.z, B2) - B2) + (long long)w2.w * (max(v3 + w.w, B3) - B3) +
           (long long)w2_5 * (max(v4 + w5, B4) - B4);
  } else {
    return (long long)w2.x * max(h.x + w.x, 0) + (long long)w2.y * max(h.y + w.y, 0) + (long long)w2.z * max(h.z + w.z, 0) +
           (long long)w2.w * max(h.w + w.w, 0);
  }
}

// NW warps per CTA, 2 CTAs per SM.  HP4 = 5 (17..20 hidden units) only: PACKED = the 24-bit state (4 int4 per site, slots of five
// values), else int32 (5 int4 per site).
// STAGES = 2: a gated step takes ~3 us per warp, far longer than an HBM round trip, so ONE window in flight per warp hides the
// latency and the shared memory of the third stage buys more warps per SM (the open kernel's step is 4 x shorter: three stages).
template <int NSLOTS, bool PACKED, int NW, int STAGES>
__global__ void __launch_bounds__(NW * 32, 2)
    sa_window_gated_kernel(SaParams P, WinGeom g, WinGate wg, const int4 *__restrict__ Wst,
                           const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_meta) {
  constexpr int HP4 = 5;
  constexpr int SI4 = PACKED ? 4 : HP4;
  constexpr int STRIDE = (32 / SI4) * SI4;
  constexpr int SPP = STRIDE / SI4;                 // sites per pass of the slot loop (8 packed, 6 int32)
  extern __shared__ __align__(128) unsigned char s_raw[];
  const int K = P.sv.K;
  const int Q = K * SI4;                            // int4 of first-layer state per window
  const int QA = (Q + 7) & ~7;                      // the records start on a 128-byte boundary of the stage (tensor-TMA destination)
  const int QM = QA + K;                            // int4 index of the stage's mbarrier: behind the records
  const int stage_i4 = g.stage_i4;
  const int tab_i4 = g.tab_i4;                      // entries of the weight tables: the state slots + zero padding
  const int lane = threadIdx.x & 31;
  const int warp = __reduce_max_sync(SO_FULL, threadIdx.x >> 5);
  const int d = P.d;
  const int Wpad = P.W + 1;
  // shared layout: [per warp: STAGES stages of (state Q | pad to 128 B | records K | the stage's mbarrier | padding)]
  // [+W][-W][zero W: tab_i4 int4 each][w2: 8 int4][per warp: bits W + 1][fifth weights +, -, zero: tab_i4 ints each (PACKED)]
  // [per warp: patch 32 words | two lists 64 x u16 | results 64 bytes][tree]
  unsigned char *s_stage = s_raw + (size_t)warp * STAGES * stage_i4 * 16;
  int4 *s_Wp = (int4 *)s_raw + (size_t)NW * STAGES * stage_i4;
  int4 *s_Wn = s_Wp + tab_i4;
  int4 *s_Wz = s_Wn + tab_i4;
  int4 *s_w2 = s_Wz + tab_i4;
  uint32_t *s_bits = (uint32_t *)(s_w2 + 8) + (size_t)warp * Wpad;
  int *s_W5p = (int *)((uint32_t *)(s_w2 + 8) + (((size_t)NW * Wpad + 3) & ~(size_t)3));
  int *s_W5n = s_W5p + (PACKED ? tab_i4 : 0);
  int *s_W5z = s_W5n + (PACKED ? tab_i4 : 0);
  uint32_t *s_patch = (uint32_t *)(s_W5z + (PACKED ? tab_i4 : 0)) + warp * 128;      // 32 patch rows, 2 x 64 list entries, 64 results
  unsigned short *s_list = (unsigned short *)(s_patch + 32), *s_swl = s_list + 64;  // rows to walk / rows that switch
  unsigned char *s_res = (unsigned char *)(s_swl + 64);
  uint2 *s_tree_buf = (uint2 *)((uint32_t *)(s_W5z + (PACKED ? tab_i4 : 0)) + NW * 128);
  const uint2 *tree = wg.tree_smem ? s_tree_buf : wg.tree;
  if (threadIdx.x < HP4) s_w2[threadIdx.x] = ((const int4 *)P.sv.w2q)[threadIdx.x];
  for (int q = threadIdx.x; q < tab_i4; q += blockDim.x) {
    int4 w = make_int4(0, 0, 0, 0);
    int w5 = 0;
    if (q < Q) {
      if constexpr (PACKED) { w = Wst[2 * q]; w5 = Wst[2 * q + 1].x; }
      else w = Wst[q];
    }
    s_Wp[q] = w;
    s_Wn[q] = make_int4(-w.x, -w.y, -w.z, -w.w);
    s_Wz[q] = make_int4(0, 0, 0, 0);
    if constexpr (PACKED) { s_W5p[q] = w5; s_W5n[q] = -w5; s_W5z[q] = 0; }
  }
  if (wg.tree_smem)
    for (int i = threadIdx.x; i < wg.n_tree; i += blockDim.x) s_tree_buf[i] = wg.tree[i];
  if (lane == 0)
    for (int i = 0; i < STAGES; ++i) mbar_init(smem_u32(s_stage) + (uint32_t)(i * stage_i4 + QM) * 16, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_async_smem();
  __syncthreads();

  const int nslots = NSLOTS ? NSLOTS : g.nslots;
  const int part = PACKED ? (lane & 3) : (lane % HP4);
  const int4 w2 = PACKED ? make_int4(P.sv.w2q[5 * part], P.sv.w2q[5 * part + 1], P.sv.w2q[5 * part + 2], P.sv.w2q[5 * part + 3])
                         : (lane < STRIDE ? s_w2[part] : make_int4(0, 0, 0, 0));
  const int w2_5 = PACKED ? P.sv.w2q[5 * part + 4] : 0;
  int B0 = 0, B1 = 0, B2 = 0, B3 = 0, B4 = 0;
  if constexpr (PACKED) {
    const int32_t *bs = P.sv.bias24 + 5 * part;
    B0 = bs[0]; B1 = bs[1]; B2 = bs[2]; B3 = bs[3]; B4 = bs[4];
  }
  // the clamped last slot must land on a zero-weight entry whose CONTENT may be anything: the padding behind the records
  const int last_off = min(lane + STRIDE * (nslots - 1), tab_i4 - 1) - lane;
  const uint32_t meta_off = (uint32_t)QA * 16, bar_off = (uint32_t)QM * 16, tx_bytes = (uint32_t)(Q + K) * 16,
                 stage_pitch = (uint32_t)stage_i4 * 16, ring_bytes = STAGES * stage_pitch;
  const uint32_t stage0 = smem_u32(s_stage);
  const float k_f = (float)(P.sv.y_scale * P.sv.c);
  const int fast_ok = !P.exact_guard && fabsf(k_f) > 1e-30f && fabsf(k_f) < 1e30f;
  const int gb1 = g.b1, gb2 = g.b2, gL = g.L, gseg = g.n_seg;
  unsigned lim1 = g.use_tma ? (unsigned)(d - gL + 1) : 0u, lim2 = g.use_tma ? (unsigned)(d - gseg + 1) : 0u;
  asm volatile("" : "+r"(lim1), "+r"(lim2));      // (opaque: otherwise recomputed from the parameters inside the step loop)
  const bool record = P.accept_out != nullptr;
  const int n_steps = (int)P.n_steps;
  const int PW = 2 * gL - 1, PH = 2 * gseg - 1;
  const bool narrow = PW <= d && PH <= d;          // the patch holds every site at most once
  // the two window positions of this lane (rounds 0 and 1 of the gate pass) and their window coordinates
  // ... packed into ONE opaque register (position | p1 << 6 | p2 << 10, for lane and lane + 32): left visible, the compiler
  // rematerialises the two divisions inside the step loop to save registers (28 instructions per step)
  unsigned ent;
  {
    const int pa2 = lane / gL, pa1 = lane - pa2 * gL;
    const int pb2 = (lane + 32) / gL, pb1 = (lane + 32) - pb2 * gL;
    ent = (unsigned)(lane | pa1 << 6 | pa2 << 10) | ((unsigned)((lane + 32) | pb1 << 6 | pb2 << 10) << 16);
    asm volatile("" : "+r"(ent));
  }
  const bool va = lane < K, vb = lane + 32 < K;
  const int site0 = lane / SI4;                     // window site of the lane's slot 0 (slot it: site0 + SPP * it)
  unsigned lt_mask;
  asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));

  uint32_t p_off = 0, par = 1;

  for (;;) {
    unsigned c32 = 0;
    if (lane == 0) c32 = (unsigned)min(atomicAdd(P.counters, 1ULL), 0xffffffffULL);
    const long long c = (long long)__reduce_max_sync(SO_FULL, c32);
    if (c >= P.n_chains) break;
    uint32_t *bits_g = P.bits + c * P.W;
    for (int w = lane; w < P.W; w += 32) s_bits[w] = bits_g[w];
    if (lane == 0) s_bits[P.W] = 0u;
    // (state and records of the chain are addressed through the tensor maps; the plain pointers are formed where the cold
    //  paths need them, not once per step)
#define SO_WG_HQ ((int4 *)P.hq + c * (long long)P.N * SI4)
#define SO_WG_META (wg.meta + c * (long long)P.N)
    long long hi = P.hi[c], lo = P.lo[c];
    int nact = P.nact[c];
    const double tsc = P.tscale[c];
    unsigned n_acc = 0, n_guard = 0;
    __syncwarp();
    if (wg.build) {      // records of the starting structure; generic-proxy stores, read by the async proxy (TMA) below
      for (int r0 = 0; r0 < P.N; r0 += 32) {
        const int r = r0 + lane;
        if (r < P.N) {
          const int r2 = r / d;
          const uint2 m = walk_site_cold(tree, s_bits, d, r - r2 * d, r2, wg.a1, wg.a2);
          SO_WG_META[r] = make_int4((int)m.x, (int)m.y, 0, 0);
        }
      }
      asm volatile("fence.proxy.async.global;" ::: "memory");
      __syncwarp();
    }

    Draw batch;
    batch.f = 0; batch.thr = 0.f;
    int f_a = 0, f_b = 0;
    float thr_a = 0.f, thr_b = 0.f;
    int hist0 = -1, hist1 = -1;
    int pend_f = -1;
    for (int s = -(STAGES - 1); s < n_steps; ++s) {
      uint32_t c_off = p_off + stage_pitch;
      if (c_off == ring_bytes) { c_off = 0; par ^= 1u; }
      // ---- 1. prefetch state + records of step s+2 ---------------------------------------------------------
      const int sp = s + (STAGES - 1);
      int f_q = 0;
      float thr_q = 0.f;
      if (sp < n_steps) {
        if ((sp & 31) == 0) {
          batch = draw_batch_cold(P.temps, P.proposals, P.uniforms, P.n_steps, P.step0, P.chain_id0 + c, c, P.seed, P.N, d,
                                  sp >> 5, lane, tsc, fast_ok);
          if (pend_f >= 0) {
            bulk_wait_all();
            pend_f = -1;
            __syncwarp();
          }
        }
        f_q = __reduce_or_sync(SO_FULL, lane == (sp & 31) ? batch.f : 0);
        thr_q = __shfl_sync(SO_FULL, batch.thr, sp & 31);
        const int q1 = pk1(f_q), q2 = pk2(f_q);
        if (pend_f >= 0 && near_cyclic(q1, pk1(pend_f), gL - 1, d) && near_cyclic(q2, pk2(pend_f), gseg - 1, d)) {
          bulk_wait_all();
          pend_f = -1;
          __syncwarp();
        }
        {
          const uint32_t dst = stage0 + p_off, bar = dst + bar_off;
          const bool interior = ((unsigned)(q1 - gb1) < lim1) & ((unsigned)(q2 - gb2) < lim2);
          if (interior) {
            if (lane == 0) {
              mbar_expect_tx(bar, tx_bytes);
              tma_load_3d(dst, &tmap, (q1 - gb1) * (SI4 * 4), q2 - gb2, (int)c, bar);
              tma_load_3d(dst + meta_off, &tmap_meta, (q1 - gb1) * 4, q2 - gb2, (int)c, bar);
            }
          } else {
            if (lane == 0) mbar_expect_tx(bar, tx_bytes);
            __syncwarp();
            boundary_loads_cold(SO_WG_HQ, dst, bar, q1, q2, gb1, gb2, gL, gseg, d, SI4, lane);
            boundary_loads_cold(SO_WG_META, dst + meta_off, bar, q1, q2, gb1, gb2, gL, gseg, d, 1, lane);
          }
        }
      }
      // ---- 2. step s ---------------------------------------------------------------------------------------
      if (s >= 0) {
        const int fp = f_b;
        const float thr = thr_b;
        const int f1 = pk1(fp), f2 = pk2(fp);
        const int f = f2 * d + f1;
        bool stale = false;
        if ((hist0 & hist1) != -1) {
          if (hist0 >= 0) stale |= near_cyclic(f1, pk1(hist0), gL - 1, d) && near_cyclic(f2, pk2(hist0), gseg - 1, d);
          if (hist1 >= 0) stale |= near_cyclic(f1, pk1(hist1), gL - 1, d) && near_cyclic(f2, pk2(hist1), gseg - 1, d);
        }
        const int xbit = (s_bits[f >> 5] >> (f & 31)) & 1;
        mbar_wait(stage0 + c_off + bar_off, par);
        int4 *stg0 = (int4 *)(s_stage + c_off);
        int4 *stg = stg0 + lane;
        if (stale) {
          bulk_wait_all();
          pend_f = -1;
          __syncwarp();
          reload_window_cold(stg0, SO_WG_HQ, f1, f2, gb1, gb2, gL, Q, d, SI4, lane);
          reload_window_cold(stg0 + QA, SO_WG_META, f1, f2, gb1, gb2, gL, K, d, 1, lane);
          __syncwarp();
        }
        // ---- gates: g0 from the records, g1 by walking the rows whose path tests f -----------------------------
        uint2 ma = make_uint2(0u, 0u), mb = make_uint2(0u, 0u);
        if (va) ma = *(const uint2 *)(stg0 + QA + lane);
        if (vb) mb = *(const uint2 *)(stg0 + QA + lane + 32);
        const unsigned ta = (ma.x >> lane) & 1u, tb = (mb.y >> lane) & 1u & (unsigned)vb;
        unsigned g1a = ma.y >> 31, g1b = mb.y >> 31;
        const unsigned walk_a = __ballot_sync(SO_FULL, ta), walk_b = __ballot_sync(SO_FULL, tb);
        const int n_walk = __popc(walk_a) + __popc(walk_b);
        if (n_walk) {
          // the rows to walk (about depth / K of the window) are compacted onto the first lanes: one round of walks
          if (lane < PH) {
            const int y = wrap(f2 - (gseg - 1) + lane, d), x0 = wrap(f1 - (gL - 1), d);
            s_patch[lane] = narrow ? patch_row_narrow(s_bits, d, y, x0, PW) ^ ((lane == gseg - 1 ? 1u : 0u) << (gL - 1))
                                   : patch_row(s_bits, d, y, x0, PW, f1, f2);
          }
          if (ta) s_list[__popc(walk_a & lt_mask)] = (unsigned short)(ent & 0xffffu);
          if (tb) s_list[__popc(walk_a) + __popc(walk_b & lt_mask)] = (unsigned short)(ent >> 16);
          __syncwarp();
          for (int base = 0; base < n_walk; base += 32) {
            if (base + lane < n_walk) {
              const int e = s_list[base + lane];      // position | p1 << 6 | p2 << 10
              const int p1 = (e >> 6) & 15, p2 = e >> 10;
              s_res[e & 63] = (unsigned char)(wg.tree_smem ? walk_gate(s_tree_buf, s_patch, p1, p2) : walk_gate(wg.tree, s_patch, p1, p2));
            }
          }
          __syncwarp();
          if (ta) g1a = s_res[lane];
          if (tb) g1b = s_res[lane + 32];
        }
        const unsigned A_lo = __ballot_sync(SO_FULL, ma.y >> 31), A_hi = __ballot_sync(SO_FULL, mb.y >> 31);
        const unsigned N_lo = __ballot_sync(SO_FULL, g1a), N_hi = __ballot_sync(SO_FULL, g1b);
        const unsigned both_lo = A_lo & N_lo, both_hi = A_hi & N_hi;
        const unsigned sw_lo = A_lo ^ N_lo, sw_hi = A_hi ^ N_hi;
        // ---- relu differences of the rows active on both sides: inactive rows read the zero weight table -------
        // the three tables (+W, -W, zero) are contiguous: one index offset selects the table of a slot
        const int4 *Wl = (xbit ? s_Wn : s_Wp) + lane;
        const int4 *Wt = s_Wp + lane;
        const int dir_off = xbit ? tab_i4 : 0, zero_off = 2 * tab_i4;
        // bit (SPP * it) of (both >> site0) <=> the site of slot `it` is active on both sides
        const unsigned m_lo = __funnelshift_r(both_lo, both_hi, site0), m_hi = both_hi >> site0;
        int Dx = 0, Dy = 0, Dz = 0, Dw = 0;
        long long acc;
        if constexpr (PACKED) {
          const int *W5t = s_W5p + lane;
          int Dv = 0;
#pragma unroll
          for (int it = 0; it < nslots; ++it) {
            const int off = (it == nslots - 1) ? last_off : STRIDE * it;
            const int sb = SPP * it;
            const bool act = sb < 32 ? ((m_lo >> sb) & 1u) : ((m_hi >> (sb - 32)) & 1u);
            const int4 h = stg[off];
            const int wi = (act ? dir_off : zero_off) + off;
            const int4 w = Wt[wi];
            const int w5 = W5t[wi];
            const int v0 = h.x & 0xffffff;
            const int v1 = (int)(__funnelshift_r((unsigned)h.x, (unsigned)h.y, 24) & 0xffffffu);
            const int v2 = (int)(__funnelshift_r((unsigned)h.y, (unsigned)h.z, 16) & 0xffffffu);
            const int v3 = (int)((unsigned)h.z >> 8);
            const int v4 = h.w;                       // (a clamped slot may read a record: any bits, its weights are zero)
            Dx += max(v0 + w.x, B0) - max(v0, B0);
            Dy += max(v1 + w.y, B1) - max(v1, B1);
            Dz += max(v2 + w.z, B2) - max(v2, B2);
            Dw += max(v3 + w.w, B3) - max(v3, B3);
            Dv += max(v4 + w5, B4) - max(v4, B4);
          }
          acc = (long long)w2.x * Dx + (long long)w2.y * Dy + (long long)w2.z * Dz + (long long)w2.w * Dw +
                (long long)w2_5 * Dv;
        } else {
#pragma unroll
          for (int it = 0; it < nslots; ++it) {
            const int off = (it == nslots - 1) ? last_off : STRIDE * it;
            const int sb = SPP * it;
            const bool act = sb < 32 ? ((m_lo >> sb) & 1u) : ((m_hi >> (sb - 32)) & 1u);
            const int4 h = stg[off];
            const int4 w = Wt[(act ? dir_off : zero_off) + off];
            Dx += max(h.x + w.x, 0) - max(h.x, 0);
            Dy += max(h.y + w.y, 0) - max(h.y, 0);
            Dz += max(h.z + w.z, 0) - max(h.z, 0);
            Dw += max(h.w + w.w, 0) - max(h.w, 0);
          }
          acc = (long long)w2.x * Dx + (long long)w2.y * Dy + (long long)w2.z * Dz + (long long)w2.w * Dw;
        }
        {
          int l0 = (int)(acc & 0x1fffff), l1 = (int)((acc >> 21) & 0x1fffff), l2 = (int)(acc >> 42);
          l0 = __reduce_add_sync(SO_FULL, l0);
          l1 = __reduce_add_sync(SO_FULL, l1);
          l2 = __reduce_add_sync(SO_FULL, l2);
          acc = ((long long)l2 << 42) + ((long long)l1 << 21) + (long long)l0;
        }
        // ---- rows whose gate switches: + the whole new row sum (on) / - the whole old one (off) ------------------
        long long dhi = acc >> SO_SPLIT_BITS, dlo = acc & SO_SPLIT_MASK;
        int dn = 0;
        const bool switching = (sw_lo | sw_hi) != 0;
        if (switching) {
          long long ph = 0;
          int pl = 0;
          // the switching rows go SPP at a time: lane group g = lane / SI4 takes the g-th of them, its lanes the row's int4 parts
          const unsigned swa = (sw_lo >> lane) & 1u, swb = (sw_hi >> lane) & 1u;
          const int n_sw = __popc(sw_lo) + __popc(sw_hi);
          if (swa) s_swl[__popc(sw_lo & lt_mask)] = (unsigned short)lane;
          if (swb) s_swl[__popc(sw_lo) + __popc(sw_hi & lt_mask)] = (unsigned short)(lane + 32);
          __syncwarp();
          for (int base = 0; base < n_sw; base += SPP) {
            const int i = base + site0;
            if (lane < STRIDE && i < n_sw) {
              const int p = s_swl[i];
              const bool on = ((p < 32 ? N_lo : N_hi) >> (p & 31)) & 1u;
              const int q = p * SI4 + part;
              const int4 h = stg0[q];
              const int wi = (on ? dir_off : zero_off) + q;
              const int4 w = s_Wp[wi];
              int w5 = 0;
              if constexpr (PACKED) w5 = s_W5p[wi];
              const long long A = row_part<PACKED>(h, w, w5, w2, w2_5, B0, B1, B2, B3, B4);
              const int sg = on ? 1 : -1;
              ph += sg * (A >> SO_SPLIT_BITS);
              pl += sg * (int)(A & SO_SPLIT_MASK);
            }
          }
          dn = __popc(N_lo & sw_lo) + __popc(N_hi & sw_hi) - __popc(A_lo & sw_lo) - __popc(A_hi & sw_hi);
          __syncwarp();
          int l0 = (int)(ph & 0x1fffff), l1 = (int)((ph >> 21) & 0x1fffff), l2 = (int)(ph >> 42);
          l0 = __reduce_add_sync(SO_FULL, l0);
          l1 = __reduce_add_sync(SO_FULL, l1);
          l2 = __reduce_add_sync(SO_FULL, l2);
          dhi += ((long long)l2 << 42) + ((long long)l1 << 21) + (long long)l0;
          dlo += (long long)__reduce_add_sync(SO_FULL, pl);    // |pl| < 2^20 * 5 lanes * K rows: fits
        }
        // ---- Metropolis: fp32 comparison with the step's threshold, fp64 rule when they are close ----------------
        bool accept;
        {
          double delta_d = 0.0;
          float x;
          if (switching) {
            delta_d = so_real(P.sv.c, P.sv.b2, P.sv.y_scale, P.sv.y_mean, dhi, dlo, dn);
            x = (float)delta_d;
          } else {
            x = (float)acc * k_f;
          }
          const float diff = x - thr;
          if (fabsf(diff) > 1e-5f * fmaxf(fabsf(x), fabsf(thr)) && fabsf(x) < 1e30f) {
            accept = diff > 0.f;
          } else {
            if (!switching) delta_d = so_real(P.sv.c, P.sv.b2, P.sv.y_scale, P.sv.y_mean, dhi, dlo, 0);
            const float u = redraw_u_cold(P.uniforms, P.n_steps, P.step0, P.chain_id0 + c, c, P.seed, s);
            int r = decide_delta_cold(delta_d, tsc, P.temps, s, u, P.exact_guard, P.guard_tol);
            if (r & 2) {
              r = exact_decision_cold(P.sv, s_bits, d, f, lane, __dmul_rn(tsc, __ldg(P.temps + s)), u);
              ++n_guard;
            }
            accept = r & 1;
          }
        }
        if (accept) {
          if (P.type_hist) so_track_flip(P.type_hist, P.n_chains, c, s_bits, d, P.N, f, lane);
          asm volatile("" ::: "memory");
#pragma unroll
          for (int it = 0; it < nslots; ++it) {
            const int q = lane + STRIDE * it;
            if (lane < STRIDE && q < Q) {
              int4 h = stg[STRIDE * it];
              const int4 w = Wl[STRIDE * it];
              if constexpr (PACKED) {
                const int w5 = ((xbit ? s_W5n : s_W5p) + lane)[STRIDE * it];
                const unsigned v0 = (unsigned)((h.x & 0xffffff) + w.x);
                const unsigned v1 = (__funnelshift_r((unsigned)h.x, (unsigned)h.y, 24) & 0xffffffu) + (unsigned)w.y;
                const unsigned v2 = (__funnelshift_r((unsigned)h.y, (unsigned)h.z, 16) & 0xffffffu) + (unsigned)w.z;
                const unsigned v3 = ((unsigned)h.z >> 8) + (unsigned)w.w;
                const unsigned v4 = (unsigned)(h.w + w5);
                h.x = (int)(v0 | (v1 << 24));
                h.y = (int)((v1 >> 8) | (v2 << 16));
                h.z = (int)((v2 >> 16) | (v3 << 8));
                h.w = (int)v4;
              } else {
                h.x += w.x; h.y += w.y; h.z += w.z; h.w += w.w;
              }
              stg[STRIDE * it] = h;
            }
          }
          // the rows that walked get their new record (path positions + gate) from a marking walk of the new structure (the
          // patch is still the flipped one); the other rows' paths do not read f
          // (the compacted list of the gate pass is still there: one round of marking walks, each writing its row's record)
          for (int base = 0; base < n_walk; base += 32) {
            if (base + lane < n_walk) {
              const int e = s_list[base + lane];
              const uint2 m = walk_mark_cold(tree, s_patch, (e >> 6) & 15, e >> 10);
              stg0[QA + (e & 63)] = make_int4((int)m.x, (int)m.y, 0, 0);
            }
          }
          fence_async_smem();
          bulk_wait_all();
          __syncwarp();
          {
            const uint32_t src = stage0 + c_off;
            const bool interior = ((unsigned)(f1 - gb1) < lim1) & ((unsigned)(f2 - gb2) < lim2);
            if (interior) {
              if (lane == 0) {
                tma_store_3d(&tmap, (f1 - gb1) * (SI4 * 4), f2 - gb2, (int)c, src);
                tma_store_3d(&tmap_meta, (f1 - gb1) * 4, f2 - gb2, (int)c, src + meta_off);
              }
            } else {
              boundary_stores_cold(SO_WG_HQ, src, f1, f2, gb1, gb2, gL, gseg, d, SI4, lane);
              boundary_stores_cold(SO_WG_META, src + meta_off, f1, f2, gb1, gb2, gL, gseg, d, 1, lane);
            }
            if (lane == 0 || !interior) {
              bulk_commit();
              asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            if (lane == 0) s_bits[f >> 5] ^= (1u << (f & 31));
          }
          pend_f = fp;
          lo += dlo;
          hi += dhi + (lo >> SO_SPLIT_BITS);
          lo &= SO_SPLIT_MASK;
          nact += dn;
          ++n_acc;
        }
        hist1 = STAGES == 3 ? hist0 : -1;            // a window is prefetched STAGES - 1 steps ahead: that many flips can outdate it
        hist0 = accept ? fp : -1;
        if (record) {
          if (lane == 0) P.accept_out[c * P.n_steps + s] = accept ? 1 : 0;
        }
        __syncwarp();
      }
      if constexpr (STAGES == 3) {
        f_b = f_a; thr_b = thr_a;
        f_a = f_q; thr_a = thr_q;
      } else {
        f_b = f_q; thr_b = thr_q;
      }
      p_off = c_off;
    }
    bulk_wait_all();
    __syncwarp();
#pragma unroll
    for (int i = 0; i < STAGES - 1; ++i) {
      if (p_off == 0) { par ^= 1u; p_off = ring_bytes; }
      p_off -= stage_pitch;
    }
    for (int w = lane; w < P.W; w += 32) bits_g[w] = s_bits[w];
    if (lane == 0) {
      P.hi[c] = hi;
      P.lo[c] = lo;
      P.nact[c] = nact;
      if (n_acc) atomicAdd(P.counters + 1, (unsigned long long)n_acc);
      if (n_guard) atomicAdd(P.counters + 2, (unsigned long long)n_guard);
    }
    __syncwarp();
  }
#undef SO_WG_HQ
#undef SO_WG_META
}

// int4 per stage: state, pad to 8, records, mbarrier, pad to 8
static int gated_stage_i4(int Q, int K) { return (((Q + 7) & ~7) + K + 1 + 7) & ~7; }

static int gated_tab_i4(int Q) { return (Q + 3 + 7) & ~7; }

constexpr int kGatedStages = 2;

static size_t gated_smem_bytes(int Q, int K, int W, int nw, bool packed, int tree_nodes) {
  const size_t st = gated_stage_i4(Q, K), tab = gated_tab_i4(Q);
  size_t b = ((size_t)nw * kGatedStages * st + 3 * tab + 8) * 16;
  b += (((size_t)nw * (W + 1) + 3) & ~(size_t)3) * 4;
  if (packed) b += 3 * tab * 4;
  b += (size_t)nw * 128 * 4;
  b += (size_t)tree_nodes * 8;
  return b + 16;
}

template <int NSLOTS, bool PACKED, int NW>
static int launch_gated(const so_chains *c, const SaParams &P, const WinGeom &g, const WinGate &wg, size_t smem, int blocks,
                        cudaStream_t st) {
  auto kern = sa_window_gated_kernel<NSLOTS, PACKED, NW, kGatedStages>;
  SO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<blocks, NW * 32, smem, st>>>(P, g, wg, (const int4 *)(PACKED ? c->sur->Wst24_dev : c->sur->Wst_dev), c->tmap,
                                      c->tmap_meta);
  SO_CUDA(cudaGetLastError());
  return SO_OK;
}

constexpr size_t kCtaSmemLimit = 113 * 1024;      // two CTAs per SM

}  // namespace

// can the gated window kernel serve this surrogate on this handle?  (17..20 hidden units, window up to 16 x 16 with at most 63
// sites, a tree with 16-bit node indices, record array and both tensor maps present)
bool so_window_gated_ok(const so_chains *c) {
  const so_surrogate *s = c->sur;
  if (!s || !s->gated || !s->win_ok || s->HP4 != 5 || !s->wtree_dev || !c->gate_meta_dev) return false;
  const int Lw = s->win_b1 - s->win_a1 + 1, nseg = s->win_b2 - s->win_a2 + 1;
  if (Lw > 16 || nseg > 16 || s->K > 63) return false;
  const int SI4 = s->compact ? 4 : 5;
  const int Q = s->K * SI4;
  const int nslots = (s->K * SI4 + (s->compact ? 32 : 30) - 1) / (s->compact ? 32 : 30);
  if (nslots > 10) return false;
  return gated_smem_bytes(Q, s->K, c->lat->W, 4, s->compact != 0, 0) <= kCtaSmemLimit;
}

int so_launch_anneal_window_gated(const so_chains *c, const SaParams &P, cudaStream_t st, int build_records) {
  const so_surrogate *s = c->sur;
  int n_sm = 0;
  SO_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, c->device));
  SO_REQUIRE(so_window_gated_ok(c), "gated window kernel: shape not supported");
  SO_REQUIRE(P.n_steps < 0x7ffffff0LL, "more than 2^31 steps in one launch: split the schedule");
  SO_REQUIRE(P.n_chains < 0xffffffffLL, "more than 2^32 chains in one handle");
  const bool packed = s->compact != 0;
  const int SI4 = packed ? 4 : 5, stride = packed ? 32 : 30;
  WinGeom g;
  g.b1 = s->win_b1; g.b2 = s->win_b2; g.L = s->win_b1 - s->win_a1 + 1; g.n_seg = s->win_b2 - s->win_a2 + 1;
  g.use_tma = c->tmap_ok && c->tmap_meta_ok;
  const int Q = s->K * SI4;
  g.stage_i4 = gated_stage_i4(Q, s->K);
  g.tab_i4 = gated_tab_i4(Q);
  g.nslots = (Q + stride - 1) / stride;
  WinGate wg;
  wg.meta = (int4 *)c->gate_meta_dev;
  wg.tree = (const uint2 *)s->wtree_dev;
  wg.n_tree = s->n_tree;
  wg.a1 = s->win_a1; wg.a2 = s->win_a2;
  wg.build = build_records ? 1 : 0;
  // warps per CTA: as many as two CTAs per SM allow (8 .. 4); the tree rides in shared memory when it still fits
  int nw = 8;
  // (9 or 10 warps per CTA fit the shared memory but not the register file: 96 registers per thread, 100+ bytes of spills,
  // 7.57e8 -> 7.07e8 / 7.25e8 flip-evaluations/s)
  while (nw > 4 && gated_smem_bytes(Q, s->K, P.W, nw, packed, 0) > kCtaSmemLimit) --nw;
  wg.tree_smem = gated_smem_bytes(Q, s->K, P.W, nw, packed, s->n_tree) <= kCtaSmemLimit ? 1 : 0;
  if (!wg.tree_smem && nw > 6 && gated_smem_bytes(Q, s->K, P.W, nw - 1, packed, s->n_tree) <= kCtaSmemLimit) {
    --nw;                                          // one warp less buys the tree a place in shared memory
    wg.tree_smem = 1;
  }
  const size_t smem = gated_smem_bytes(Q, s->K, P.W, nw, packed, wg.tree_smem ? s->n_tree : 0);
  long long want = (P.n_chains + nw - 1) / nw;
  int blocks = (int)(want < (long long)n_sm * 2 ? want : (long long)n_sm * 2);
  if (blocks < 1) blocks = 1;
#define SO_WG_CASE(NWV)                                                                                        \
  case NWV:                                                                                                    \
    if (packed) {                                                                                              \
      if (g.nslots == 7) return launch_gated<7, true, NWV>(c, P, g, wg, smem, blocks, st);                     \
      return launch_gated<0, true, NWV>(c, P, g, wg, smem, blocks, st);                                        \
    }                                                                                                          \
    if (g.nslots == 9) return launch_gated<9, false, NWV>(c, P, g, wg, smem, blocks, st);                      \
    return launch_gated<0, false, NWV>(c, P, g, wg, smem, blocks, st);
  switch (nw) {
    SO_WG_CASE(8) SO_WG_CASE(7) SO_WG_CASE(6) SO_WG_CASE(5) SO_WG_CASE(4)
  }
#undef SO_WG_CASE
  so_set_error("gated window kernel: no launch configuration");
  return SO_EINVAL;
}
