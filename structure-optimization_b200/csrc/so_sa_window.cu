// so_sa_window.cu -- the production SA kernel for rectangular-window descriptors (e.g. the 7x7 local window of
// OML/dynamic_cat.py:325-341), gate off.  Same arithmetic and decisions as sa_flat_kernel (so_sa.cu) -- exact
// integer sums, fp64 Metropolis -- restructured for the B200 memory system:
//
//   * one warp per chain; the 4*K*Hp bytes of first-layer state a flip touches are (2*R2+1) contiguous
//     segments of the chain's hq array.  They are fetched by bulk-async copies (cp.async.bulk, the non-tensor TMA
//     path, SASS UBLKCP) into a per-warp ring of shared-memory stages, completion on an mbarrier per stage;
//   * proposals are known ahead of time (counter-based Philox), so the copies for step s+STAGES-1 are issued while
//     step s computes: STAGES-1 windows per warp are in flight and the HBM latency is off the critical path;
//   * an accepted flip writes the updated window back with one tensor-TMA store from the same stage buffer.
//     Rare hazards (the window of a prefetched step overlaps a flip accepted after the prefetch was issued, or a
//     store still in flight) are detected exactly and resolved by draining the stores and re-fetching;
//   * Philox and the injected-sequence loads are evaluated lane-parallel for 32 steps at a time; the warp sum uses
//     three REDUX.SUM on 21-bit limbs; the decision is an fp32 comparison with a per-draw threshold thr = min(T ln u, 0)
//     computed in fp64 with the draws, the fp64 rule runs only when the two sides are within 1e-5 of each other.
#include "so_window_common.cuh"
#include "so_types.cuh"
#include <stdlib.h>

namespace {

constexpr int WARPS = 8;          // warps (= chains in flight) per CTA of the int32-state kernel; the packed one may take more

// slow, exact side of the Metropolis rule for one step; reloads the fp64 temperature and redraws the step's uniform (the hot
// loop carries only the fp32 threshold).  bit 0 accept, bit 1 guard
__device__ __noinline__ int decide_cold(long long acc, double k_sigma, double k_c, double tsc, const double *temps,
                                        long long s, float u, int exact_guard, double tol) {
  // == so_real(c, b2, sigma, mu, acc >> 20, acc & mask, 0): with n = 0 the pinned sequence reduces to
  // sigma * (c * RN(acc)) (the dropped terms add +0.0; |acc| < 2^62 so hi*2^20 + lo rounds once, like the cast)
  const double delta = __dmul_rn(k_sigma, __dmul_rn(k_c, __ll2double_rn(acc)));
  const double T = __dmul_rn(tsc, __ldg(temps + s));
  return metropolis_cold(delta, T, u, exact_guard, tol);
}

// COMMIT selects how an accepted flip is written back: 0 = updated in the shared-memory stage + one tensor-TMA store
// (production); 1 = straight from registers with st.global.cg + fence.proxy.async.global, nothing ever waits on the write
// (SO_WIN_COMMIT=1); 2 = as 1 without the proxy fence (timing experiment only).  Measured on one box, 65,536 chains
// (profiles/r02_commit_ab.txt): at accept fractions 0.93 / 0.54 / 0.21 the TMA store takes 954 / 711 / 521 ms per sweep,
// the register store 1001 / 769 / 576 ms, with or without the fence -- the write path is not latency bound, the nine
// STG.128 + address arithmetic per accepted flip cost more issue slots than one UTMASTG.
//
// PACKED (so_surrogate_create_compact): the state is stored as 24-bit biased values, u = hq + bias_j in [0, 2^24) with
// bias_j = -(smallest reachable hq of unit j) (the whole 24 bits cover the unit's L1 range, the intercept costs nothing), five
// to an int4 (15 bytes + 1 pad): a site of 20 hidden units is FOUR int4 = 64 bytes instead of 80.  Records are then aligned
// with the 64-byte L2 fill granule (the 1.13 x over-fetch of the 80-byte record disappears, profiles/
// r02_l2_fetch_granularity.txt) and a flip moves 3,136 instead of 3,920 bytes.  Same integers, same decisions: the first
// layer sits on a grid ~2^6 coarser (sum_k |W1q_kj| < 2^24), which the oracle mirrors (oracle/surrogate.py quantize(bits='packed24')).
// A lane owns int4 slots q = lane + 32*it: site q >> 2, part q & 3 = lane & 3 -> hidden units 5*part .. 5*part+4.
template <int HP4, int NSLOTS, int COMMIT, bool PACKED, int NW = WARPS>
__global__ void __launch_bounds__(NW * 32, 2)
    sa_window_kernel(SaParams P, WinGeom g, const int4 *__restrict__ Wst, const __grid_constant__ CUtensorMap tmap) {
  constexpr int STAGES = 3;
  constexpr int SI4 = PACKED ? 4 : HP4;           // int4 per site in memory
  // lanes 0..STRIDE-1 own int4 slots q = lane + STRIDE*it, it < NSLOTS; STRIDE is a multiple of SI4, so a lane always
  // sees the same hidden units and its second-layer weights live in registers
  constexpr int STRIDE = (32 / SI4) * SI4;
  extern __shared__ __align__(128) unsigned char s_raw[];
  const int Q = P.sv.K * SI4;                     // int4 per window
  const int stage_i4 = g.stage_i4;                // window + zero-weight padding, multiple of 8 int4
  const int lane = threadIdx.x & 31;
  const int warp = __reduce_max_sync(SO_FULL, threadIdx.x >> 5);   // (uniform to the compiler: addresses derived from it stay in uniform registers)
  const int d = P.d;
  // shared layout: [per warp: STAGES stages][+W: stage_i4][-W: stage_i4][w2: 8 int4][per warp: bits][PACKED: fifth weights]
  // The mbarrier of a stage lives INSIDE the stage, in the first zero-weight padding entry behind the window (int4 index Q):
  // one running byte offset addresses the data and the barrier of the stage in use
  unsigned char *s_stage = s_raw + (size_t)warp * STAGES * stage_i4 * 16;
  int4 *s_Wp = (int4 *)s_raw + (size_t)NW * STAGES * stage_i4;
  int4 *s_Wn = s_Wp + stage_i4;
  int4 *s_w2 = s_Wn + stage_i4;
  uint32_t *s_bits = (uint32_t *)(s_w2 + 8) + (size_t)warp * P.W;
  // PACKED: the fifth weight of every slot, + and - tables, after the per-warp words (stage_i4 ints each)
  int *s_W5p = (int *)((uint32_t *)(s_w2 + 8) + (((size_t)NW * P.W + 3) & ~(size_t)3));
  int *s_W5n = s_W5p + stage_i4;
  if (threadIdx.x < HP4) s_w2[threadIdx.x] = ((const int4 *)P.sv.w2q)[threadIdx.x];
  // (entries beyond the window are zero: the clamped last slot and the padding contribute exactly 0)
  if constexpr (PACKED) {
    // Wst = [Q slots][8 int32]: the five first-layer weights of the slot's hidden units, then padding
    for (int q = threadIdx.x; q < stage_i4; q += blockDim.x) {
      int4 w = make_int4(0, 0, 0, 0);
      int w5 = 0;
      if (q < Q) { w = Wst[2 * q]; w5 = Wst[2 * q + 1].x; }
      s_Wp[q] = w;
      s_Wn[q] = make_int4(-w.x, -w.y, -w.z, -w.w);
      s_W5p[q] = w5;
      s_W5n[q] = -w5;
    }
  } else {
    for (int q = threadIdx.x; q < stage_i4; q += blockDim.x) {
      int4 w = q < Q ? Wst[q] : make_int4(0, 0, 0, 0);
      s_Wp[q] = w;
      s_Wn[q] = make_int4(-w.x, -w.y, -w.z, -w.w);
    }
  }
  if (lane == 0)
    for (int i = 0; i < STAGES; ++i) mbar_init(smem_u32(s_stage) + (uint32_t)(i * stage_i4 + Q) * 16, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_async_smem();
  __syncthreads();

  // NSLOTS == 0: generic instantiation (slot count from the geometry, loops not unrolled) for shapes other than the
  // 7x7 / 20-unit headline shape
  const int nslots = NSLOTS ? NSLOTS : g.nslots;
  const int4 w2 = PACKED ? make_int4(P.sv.w2q[5 * (lane & 3)], P.sv.w2q[5 * (lane & 3) + 1], P.sv.w2q[5 * (lane & 3) + 2],
                                     P.sv.w2q[5 * (lane & 3) + 3])
                         : (lane < STRIDE ? s_w2[lane % HP4] : make_int4(0, 0, 0, 0));   // idle lanes re-read data with weight 0
  const int w2_5 = PACKED ? P.sv.w2q[5 * (lane & 3) + 4] : 0;
  // PACKED: stored value u = hq + bias_j with bias_j = -(smallest reachable hq of unit j), so relu(hq) = max(u, bias_j) - bias_j
  int B0 = 0, B1 = 0, B2 = 0, B3 = 0, B4 = 0;
  if constexpr (PACKED) {
    const int32_t *bs = P.sv.bias24 + 5 * (lane & 3);
    B0 = bs[0]; B1 = bs[1]; B2 = bs[2]; B3 = bs[3]; B4 = bs[4];
  }
  // every slot but the last stays inside the padded stage; the last one is clamped onto a zero-weight padding entry
  const int last_off = min(lane + STRIDE * (nslots - 1), stage_i4 - 1) - lane;
  const uint32_t stage_bytes = (uint32_t)Q * 16, stage_pitch = (uint32_t)stage_i4 * 16, ring_bytes = STAGES * stage_pitch;
  const uint32_t stage0 = smem_u32(s_stage);
  // fp32 image of delta = sigma * c * acc (signed scale).  A scale that is not a normal fp32 number disables the fp32 side
  const float k_f = (float)(P.sv.y_scale * P.sv.c);
  const int fast_ok = !P.exact_guard && fabsf(k_f) > 1e-30f && fabsf(k_f) < 1e30f;
  const int gb1 = g.b1, gb2 = g.b2, gL = g.L, gseg = g.n_seg;
  // a window is interior (one tensor-TMA box) iff 0 <= f1 - b1 <= d - L and 0 <= f2 - b2 <= d - n_seg: two unsigned tests
  const unsigned lim1 = g.use_tma ? (unsigned)(d - gL + 1) : 0u, lim2 = g.use_tma ? (unsigned)(d - gseg + 1) : 0u;
  const bool record = P.accept_out != nullptr;
  const int n_steps = (int)P.n_steps;               // (the launcher refuses launches of 2^31 steps or more)

  // ring of stages, continuous over all chains of the warp: the k-th window the warp ever computes on sits in stage k % 3
  // and was the k-th load issued, so its mbarrier phase has parity (k / 3) & 1.  Iteration s of a chain prefetches into
  // the stage at byte offset p_off (the one step s-1 computed in) and computes in the next stage round the ring (c_off);
  // par flips whenever c_off wraps.  (p_off, par) = the state after computing window k-1; start: k = -2.
  uint32_t p_off = 0, par = 1;

  for (;;) {
    unsigned c32 = 0;
    if (lane == 0) c32 = (unsigned)min(atomicAdd(P.counters, 1ULL), 0xffffffffULL);
    const long long c = (long long)__reduce_max_sync(SO_FULL, c32);     // (uniform; the launcher refuses 2^32 chains or more)
    if (c >= P.n_chains) break;
    uint32_t *bits_g = P.bits + c * P.W;
    for (int w = lane; w < P.W; w += 32) s_bits[w] = bits_g[w];
    int4 *hq = (int4 *)P.hq + c * (long long)P.N * SI4;
    long long hi = P.hi[c], lo = P.lo[c];
    const double tsc = P.tscale[c];
    unsigned n_acc = 0, n_guard = 0;
    __syncwarp();

    Draw batch;                                     // lane l: draws of step 32*b + l
    batch.f = 0; batch.thr = 0.f;
    // pipeline registers: draws of the steps whose windows are in flight
    int f_a = 0, f_b = 0;
    float thr_a = 0.f, thr_b = 0.f;
    int hist0 = -1, hist1 = -1;                     // flips accepted at steps s-1, s-2 (their windows were prefetched earlier)
    int pend_f = -1;                                // COMMIT == 0: flip whose TMA store may still be in flight
    for (int s = -(STAGES - 1); s < n_steps; ++s) {
      uint32_t c_off = p_off + stage_pitch;
      if (c_off == ring_bytes) { c_off = 0; par ^= 1u; }
      // ---- 1. prefetch the window of step s+2 into the stage freed by step s-1 -------------------------------
      const int sp = s + (STAGES - 1);
      int f_q = 0;
      float thr_q = 0.f;
      if (sp < n_steps) {
        if ((sp & 31) == 0) {
          batch = draw_batch_cold(P.temps, P.proposals, P.uniforms, P.n_steps, P.step0, P.chain_id0 + c, c, P.seed, P.N, d,
                                  sp >> 5, lane, tsc, fast_ok);
          // a store issued at least two steps ago has long landed: retire it here, once per 32 steps, so that the per-step
          // overlap test below runs only right after an accepted flip
          if (COMMIT == 0 && pend_f >= 0) {
            bulk_wait_all();
            pend_f = -1;
            __syncwarp();
          }
        }
        f_q = __reduce_or_sync(SO_FULL, lane == (sp & 31) ? batch.f : 0);   // REDUX: the compiler knows the result is warp-uniform
        thr_q = __shfl_sync(SO_FULL, batch.thr, sp & 31);
        const int q1 = pk1(f_q), q2 = pk2(f_q);
        if (COMMIT == 0 && pend_f >= 0 && near_cyclic(q1, pk1(pend_f), gL - 1, d) && near_cyclic(q2, pk2(pend_f), gseg - 1, d)) {
          bulk_wait_all();                          // the store of that region may still be in flight
          pend_f = -1;
          __syncwarp();
        }
        {
          const uint32_t dst = stage0 + p_off, bar = dst + stage_bytes;
          const bool interior = ((unsigned)(q1 - gb1) < lim1) & ((unsigned)(q2 - gb2) < lim2);
          if (interior) {
            if (lane == 0) {
              mbar_expect_tx(bar, stage_bytes);
              tma_load_3d(dst, &tmap, (q1 - gb1) * (SI4 * 4), q2 - gb2, (int)c, bar);
            }
          } else {                                  // warp-uniform branch; the whole warp enters the call converged
            if (lane == 0) mbar_expect_tx(bar, stage_bytes);
            __syncwarp();
            boundary_loads_cold(hq, dst, bar, q1, q2, gb1, gb2, gL, gseg, d, SI4, lane);
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
        mbar_wait(stage0 + c_off + stage_bytes, par);
        int4 *stg = (int4 *)(s_stage + c_off) + lane;
        if (stale) {                                // the writers' stores are ordered before this by their __syncwarp
          if (COMMIT == 0) {
            bulk_wait_all();                        // every lane: the stores it issued are complete and visible
            pend_f = -1;
            __syncwarp();
          }
          reload_window_cold(stg - lane, hq, f1, f2, gb1, gb2, gL, Q, d, SI4, lane);
          __syncwarp();
        }
        // delta-sum: D_j = sum over this lane's slots of relu(h + dw) - relu(h), dw = +-W (table chosen by the flip
        // direction).  |D_j| <= sum_k |W1q[k][j]| < 2^31 by construction of the fixed-point grid, so int32 is exact.
        const int4 *Wl = (xbit ? s_Wn : s_Wp) + lane;
        int Dx = 0, Dy = 0, Dz = 0, Dw = 0;
        long long acc;
        if constexpr (PACKED) {
          const int *W5 = (xbit ? s_W5n : s_W5p) + lane;
          int Dv = 0;
#pragma unroll
          for (int it = 0; it < nslots; ++it) {
            const int off = (it == nslots - 1) ? last_off : STRIDE * it;
            const int4 h = stg[off];
            const int4 w = Wl[off];
            const int w5 = W5[off];
            // five 24-bit biased values in 15 bytes (the pad byte is zero); relu(hq) = max(u, bias) - bias, the bias
            // cancels in the difference.  A padding entry may hold an mbarrier: any bits, its weights are zero
            const int v0 = h.x & 0xffffff;
            const int v1 = (int)(__funnelshift_r((unsigned)h.x, (unsigned)h.y, 24) & 0xffffffu);
            const int v2 = (int)(__funnelshift_r((unsigned)h.y, (unsigned)h.z, 16) & 0xffffffu);
            const int v3 = (int)((unsigned)h.z >> 8);
            const int v4 = h.w;
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
            const int4 h = stg[off];
            const int4 w = Wl[off];
            Dx += max(h.x + w.x, 0) - max(h.x, 0);
            Dy += max(h.y + w.y, 0) - max(h.y, 0);
            Dz += max(h.z + w.z, 0) - max(h.z, 0);
            Dw += max(h.w + w.w, 0) - max(h.w, 0);
          }
          acc = (long long)w2.x * Dx + (long long)w2.y * Dy + (long long)w2.z * Dz + (long long)w2.w * Dw;
        }
        // exact warp sum of int64 via three 21-bit limbs (each limb sum fits 32 bits)
        {
          int l0 = (int)(acc & 0x1fffff), l1 = (int)((acc >> 21) & 0x1fffff), l2 = (int)(acc >> 42);
          l0 = __reduce_add_sync(SO_FULL, l0);
          l1 = __reduce_add_sync(SO_FULL, l1);
          l2 = __reduce_add_sync(SO_FULL, l2);
          acc = ((long long)l2 << 42) + ((long long)l1 << 21) + (long long)l0;
        }
        // Metropolis (OML/optimize_SA.py:108-116) as "delta > thr" (draw_batch_cold): the fp32 comparison decides unless
        // the two sides are within 1e-5 (relative) of each other -- then the fp64 rule (decide_cold) with the redrawn u
        bool accept;
        {
          const float x = (float)acc * k_f;
          const float diff = x - thr;
          if (fabsf(diff) > 1e-5f * fmaxf(fabsf(x), fabsf(thr))) {
            accept = diff > 0.f;
          } else {
            const float u = redraw_u_cold(P.uniforms, P.n_steps, P.step0, P.chain_id0 + c, c, P.seed, s);
            int r = decide_cold(acc, P.sv.y_scale, P.sv.c, tsc, P.temps, s, u, P.exact_guard, P.guard_tol);
            if (r & 2) {
              r = exact_decision_cold(P.sv, s_bits, d, f, lane, __dmul_rn(tsc, __ldg(P.temps + s)), u);
              ++n_guard;
            }
            accept = r & 1;
          }
        }
        if (accept) {
          if (P.type_hist) so_track_flip(P.type_hist, P.n_chains, c, s_bits, d, P.N, f, lane);   // before the bit flips
          if constexpr (COMMIT == 0) {
            asm volatile("" ::: "memory");            // re-read the stage rather than keeping updated values live
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
            fence_async_smem();                       // generic-proxy writes -> visible to the bulk (async proxy) store
            bulk_wait_all();                          // at most one flip's stores in flight (every lane waits for its own groups)
            __syncwarp();
            {
              const uint32_t src = stage0 + c_off;
              const bool interior = ((unsigned)(f1 - gb1) < lim1) & ((unsigned)(f2 - gb2) < lim2);
              if (interior) {
                if (lane == 0) tma_store_3d(&tmap, (f1 - gb1) * (SI4 * 4), f2 - gb2, (int)c, src);
              } else {
                boundary_stores_cold(hq, src, f1, f2, gb1, gb2, gL, gseg, d, SI4, lane);     // warp-uniform branch
              }
              if (lane == 0 || !interior) {
                bulk_commit();
                // the stage is re-used by the prefetch issued at the next step: the store must have read it by then
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              }
              if (lane == 0) s_bits[f >> 5] ^= (1u << (f & 31));
            }
            pend_f = fp;
          } else {
            // commit: h + dw of every slot goes from registers to the chain's state in global memory
            {
              const int part = lane % HP4, pbase = lane / HP4;
#pragma unroll
              for (int it = 0; it < nslots; ++it) {
                const int q = lane + STRIDE * it;
                if (lane < STRIDE && q < Q) {
                  int4 h = stg[STRIDE * it];
                  const int4 w = Wl[STRIDE * it];
                  h.x += w.x; h.y += w.y; h.z += w.z; h.w += w.w;
                  const int p = pbase + (STRIDE / HP4) * it;          // site of the window, row-major over (p2, p1)
                  const int p2 = p / gL, p1 = p - p2 * gL;
                  int r1 = f1 - gb1 + p1, r2 = f2 - gb2 + p2;
                  r1 += r1 < 0 ? d : 0; r1 -= r1 >= d ? d : 0;
                  r2 += r2 < 0 ? d : 0; r2 -= r2 >= d ? d : 0;
                  __stcg(hq + (size_t)(r2 * d + r1) * HP4 + part, h);
                }
              }
            }
            // generic-proxy stores -> later async-proxy (TMA) reads of the same bytes: each writer fences, the
            // __syncwarp at the end of the step orders lane 0's next prefetch after all of them
            if (COMMIT == 1) asm volatile("fence.proxy.async.global;" ::: "memory");
            if (lane == 0) s_bits[f >> 5] ^= (1u << (f & 31));
          }
          lo += acc & SO_SPLIT_MASK;
          hi += (acc >> SO_SPLIT_BITS) + (lo >> SO_SPLIT_BITS);
          lo &= SO_SPLIT_MASK;
          ++n_acc;
        }
        hist1 = hist0;
        hist0 = accept ? fp : -1;
        if (record) {
          if (lane == 0) P.accept_out[c * P.n_steps + s] = accept ? 1 : 0;
        }
        __syncwarp();
      }
      f_b = f_a; thr_b = thr_a;
      f_a = f_q; thr_a = thr_q;
      p_off = c_off;
    }
    if (COMMIT == 0) bulk_wait_all();
    __syncwarp();
    // the first two iterations of the next chain compute nothing: step the ring state back by two windows
#pragma unroll
    for (int i = 0; i < STAGES - 1; ++i) {
      if (p_off == 0) { par ^= 1u; p_off = ring_bytes; }
      p_off -= stage_pitch;
    }
    for (int w = lane; w < P.W; w += 32) bits_g[w] = s_bits[w];
    if (lane == 0) {
      P.hi[c] = hi;
      P.lo[c] = lo;
      if (n_acc) atomicAdd(P.counters + 1, (unsigned long long)n_acc);
      if (n_guard) atomicAdd(P.counters + 2, (unsigned long long)n_guard);
    }
    __syncwarp();
  }
}

static int window_stage_i4(int Q) { return (Q + 3 + 7) & ~7; }

static size_t window_smem_bytes(int Q, int W, int stages, int nw = WARPS) {
  size_t st_i4 = window_stage_i4(Q);
  size_t i4 = 8 + 2 * st_i4 + (size_t)nw * stages * st_i4;
  size_t words = ((size_t)nw * W + 3) & ~(size_t)3;
  return i4 * 16 + words * 4 + 16 + 2 * st_i4 * 4;   // (+ the fifth-weight tables of PACKED)
}

// int32 state [sites][20] -> 24-bit biased, five values per int4, four int4 (64 bytes) per site
__global__ void pack24_kernel(const int4 *__restrict__ src, int4 *__restrict__ dst, long long n_sites,
                              const int32_t *__restrict__ bias) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n_sites * 4) return;
  const long long site = t >> 2;
  const int part = (int)(t & 3);
  const int *h = (const int *)(src + site * 5) + 5 * part;
  const int32_t *b = bias + 5 * part;
  const unsigned v0 = (unsigned)(h[0] + b[0]) & 0xffffffu, v1 = (unsigned)(h[1] + b[1]) & 0xffffffu,
                 v2 = (unsigned)(h[2] + b[2]) & 0xffffffu, v3 = (unsigned)(h[3] + b[3]) & 0xffffffu,
                 v4 = (unsigned)(h[4] + b[4]) & 0xffffffu;
  dst[t] = make_int4((int)(v0 | (v1 << 24)), (int)((v1 >> 8) | (v2 << 16)), (int)((v2 >> 16) | (v3 << 8)), (int)v4);
}
__global__ void unpack24_kernel(const int4 *__restrict__ src, int *__restrict__ dst, long long n_sites,
                                const int32_t *__restrict__ bias) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n_sites * 4) return;
  const int4 h = src[t];
  int *o = dst + (t >> 2) * 20 + 5 * (int)(t & 3);
  const int32_t *b = bias + 5 * (int)(t & 3);
  o[0] = (int)((unsigned)h.x & 0xffffffu) - b[0];
  o[1] = (int)(__funnelshift_r((unsigned)h.x, (unsigned)h.y, 24) & 0xffffffu) - b[1];
  o[2] = (int)(__funnelshift_r((unsigned)h.y, (unsigned)h.z, 16) & 0xffffffu) - b[2];
  o[3] = (int)((unsigned)h.z >> 8) - b[3];
  o[4] = (int)((unsigned)h.w & 0xffffffu) - b[4];
}

}  // namespace

size_t so_window_smem_bytes(int Q, int W, int stages) { return window_smem_bytes(Q, W, stages); }

int so_launch_pack24(const int32_t *src_dev, void *dst_dev, int64_t n_sites, const int32_t *bias_dev, cudaStream_t st) {
  if (n_sites == 0) return SO_OK;
  pack24_kernel<<<(unsigned)((n_sites * 4 + 255) / 256), 256, 0, st>>>((const int4 *)src_dev, (int4 *)dst_dev, n_sites, bias_dev);
  SO_CUDA(cudaGetLastError());
  return SO_OK;
}
int so_launch_unpack24(const void *src_dev, int32_t *dst_dev, int64_t n_sites, const int32_t *bias_dev, cudaStream_t st) {
  if (n_sites == 0) return SO_OK;
  unpack24_kernel<<<(unsigned)((n_sites * 4 + 255) / 256), 256, 0, st>>>((const int4 *)src_dev, dst_dev, n_sites, bias_dev);
  SO_CUDA(cudaGetLastError());
  return SO_OK;
}

namespace {

template <int HP4, int NSLOTS, int COMMIT, bool PACKED, int NW = WARPS>
static int launch_window(const so_chains *c, const SaParams &P, const WinGeom &g, size_t smem, int blocks, cudaStream_t st) {
  auto kern = sa_window_kernel<HP4, NSLOTS, COMMIT, PACKED, NW>;
  SO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<blocks, NW * 32, smem, st>>>(P, g, (const int4 *)(PACKED ? c->sur->Wst24_dev : c->sur->Wst_dev), c->tmap);
  SO_CUDA(cudaGetLastError());
  return SO_OK;
}

static int window_commit_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char *e = getenv("SO_WIN_COMMIT");
    mode = e ? atoi(e) : 0;
    if (mode < 0 || mode > 2) mode = 0;
  }
  return mode;
}

template <int HP4, int NSLOTS>
static int launch_window_commit(const so_chains *c, const SaParams &P, const WinGeom &g, size_t smem, int blocks,
                                cudaStream_t st) {
  switch (window_commit_mode()) {
    case 2: return launch_window<HP4, NSLOTS, 2, false>(c, P, g, smem, blocks, st);
    case 1: return launch_window<HP4, NSLOTS, 1, false>(c, P, g, smem, blocks, st);
    default: return launch_window<HP4, NSLOTS, 0, false>(c, P, g, smem, blocks, st);
  }
}

// the shape every BASELINE configuration with a local descriptor uses (7x7 window, 17..20 hidden units) is fully
// unrolled; any other window / width runs the generic instantiation of its HP4
template <int HP4>
static int launch_window_hp4(const so_chains *c, const SaParams &P, const WinGeom &g, size_t smem, int blocks,
                             cudaStream_t st) {
  if (g.nslots > 10) { so_set_error("window too large for the pipelined kernel"); return SO_EINVAL; }
  if (HP4 == 5 && g.nslots == 9) return launch_window_commit<5, 9>(c, P, g, smem, blocks, st);
  return launch_window<HP4, 0, 0, false>(c, P, g, smem, blocks, st);
}

}  // namespace

int so_launch_anneal_window(const so_chains *c, const SaParams &P, cudaStream_t st) {
  const so_surrogate *s = c->sur;
  int n_sm = 0;
  SO_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, c->device));
  WinGeom g;
  g.b1 = s->win_b1; g.b2 = s->win_b2; g.L = s->win_b1 - s->win_a1 + 1; g.n_seg = s->win_b2 - s->win_a2 + 1;
  g.use_tma = c->tmap_ok;
  g.tab_i4 = 0;
  long long want = (P.n_chains + WARPS - 1) / WARPS;
  int blocks = (int)(want < (long long)n_sm * 2 ? want : (long long)n_sm * 2);
  if (blocks < 1) blocks = 1;
  SO_REQUIRE(P.n_steps < 0x7ffffff0LL, "more than 2^31 steps in one launch: split the schedule");
  SO_REQUIRE(P.n_chains < 0xffffffffLL, "more than 2^32 chains in one handle");
  if (s->compact) {                              // 24-bit packed state: 4 int4 per site, all 32 lanes own slots
    const int Q = s->K * 4;
    g.stage_i4 = window_stage_i4(Q);
    g.nslots = (Q + 31) / 32;
    if (g.nslots > 10) { so_set_error("window too large for the pipelined kernel"); return SO_EINVAL; }
    // (10 or 12 warps per CTA fit the smaller packed stages, but at 96 / 80 registers per thread the spills cost more than
    // the extra warps hide: 1.44e9 -> 1.12e9 flip-evaluations/s on one box)
    if (g.nslots == 7) return launch_window<5, 7, 0, true>(c, P, g, window_smem_bytes(Q, P.W, 3), blocks, st);
    return launch_window<5, 0, 0, true>(c, P, g, window_smem_bytes(Q, P.W, 3), blocks, st);
  }
  const int Q = s->K * s->HP4;
  const int stride = (32 / s->HP4) * s->HP4;
  g.stage_i4 = window_stage_i4(Q);
  g.nslots = (Q + stride - 1) / stride;
  size_t smem = so_window_smem_bytes(Q, P.W, 3);
  switch (s->HP4) {
    case 1: return launch_window_hp4<1>(c, P, g, smem, blocks, st);
    case 2: return launch_window_hp4<2>(c, P, g, smem, blocks, st);
    case 3: return launch_window_hp4<3>(c, P, g, smem, blocks, st);
    case 4: return launch_window_hp4<4>(c, P, g, smem, blocks, st);
    case 5: return launch_window_hp4<5>(c, P, g, smem, blocks, st);
    case 6: return launch_window_hp4<6>(c, P, g, smem, blocks, st);
    case 7: return launch_window_hp4<7>(c, P, g, smem, blocks, st);
    case 8: return launch_window_hp4<8>(c, P, g, smem, blocks, st);
    default: so_set_error("hidden width not supported"); return SO_EINVAL;
  }
}
