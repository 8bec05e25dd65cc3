// so_sa_window.cu -- the production SA kernel for rectangular-window descriptors (e.g. the 7x7 local window of
// OML/dynamic_cat.py:325-341), gate off.  Same arithmetic and decisions as sa_flat_kernel (so_sa.cu) -- exact
// integer sums, fp64 Metropolis -- restructured for the B200 memory system:
//
//   * one warp per chain; the 4*K*Hp bytes of first-layer state a flip touches are (2*R2+1) contiguous
//     segments of the chain's hq array.  They are fetched by bulk-async copies (cp.async.bulk, the non-tensor TMA
//     path, SASS UBLKCP) into a per-warp ring of shared-memory stages, completion on an mbarrier per stage;
//   * proposals are known ahead of time (counter-based Philox), so the copies for step s+STAGES-1 are issued while
//     step s computes: STAGES-1 windows per warp are in flight and the HBM latency is off the critical path;
//   * an accepted flip writes the updated window back with bulk-async stores from the same stage buffer.
//     Rare hazards (the window of a prefetched step overlaps a flip accepted after the prefetch was issued, or a
//     store still in flight) are detected exactly and resolved by draining the stores and re-fetching;
//   * Philox and the injected-sequence loads are evaluated lane-parallel for 32 steps at a time; the warp sum uses
//     three REDUX.SUM on 21-bit limbs; exp() runs in fp64 only when an fp32 pre-filter cannot decide.
#include "so_internal.cuh"
#include <math.h>

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct WinGeom {
  int b1, b2;        // largest offsets: window row p2 / column p1 <-> site (f1 - b1 + p1, f2 - b2 + p2)
  int L, n_seg;      // sites per segment, number of segments
};

// cyclic |a - b| <= r
__device__ __forceinline__ bool near_cyclic(int a, int b, int r, int d) {
  int t = a - b;
  t = t < 0 ? -t : t;
  t = min(t, d - t);
  return t <= r;
}

constexpr int WARPS = 8;

template <int HP4, int STAGES>
__global__ void __launch_bounds__(WARPS * 32, 2) sa_window_kernel(SaParams P, WinGeom g, const int4 *__restrict__ Wst) {
  constexpr int MAXIT = 8;
  extern __shared__ __align__(128) unsigned char s_raw[];
  const SurView &sv = P.sv;
  const int Q = sv.K * HP4;                       // int4 per window
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = P.d;
  // shared layout: [w2: HP4 int4][per warp: STAGES stages of Q int4][per warp: bits W u32][per warp: STAGES mbarriers]
  int4 *s_w2 = (int4 *)s_raw;
  const int stage_i4 = Q;                                                     // int4 per stage
  int4 *s_stage = s_w2 + 8 + (size_t)warp * STAGES * stage_i4;                // 8 int4 reserved for w2
  uint32_t *s_bits = (uint32_t *)(s_w2 + 8 + (size_t)WARPS * STAGES * stage_i4) + (size_t)warp * P.W;
  unsigned long long *s_bar =
      (unsigned long long *)((uint32_t *)(s_w2 + 8 + (size_t)WARPS * STAGES * stage_i4) + (((size_t)WARPS * P.W + 3) & ~(size_t)3)) +
      warp * STAGES;
  if (threadIdx.x < HP4) s_w2[threadIdx.x] = ((const int4 *)sv.w2q)[threadIdx.x];
  if (lane == 0)
    for (int i = 0; i < STAGES; ++i) mbar_init(smem_u32(s_bar + i), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_async_smem();
  __syncthreads();

  // lane-constant tables: this lane's int4 slots q = lane + 32*it of the window, their weights
  int4 Wreg[MAXIT];
  int part[MAXIT];
#pragma unroll
  for (int it = 0; it < MAXIT; ++it) {
    int q = lane + 32 * it;
    Wreg[it] = q < Q ? Wst[q] : make_int4(0, 0, 0, 0);
    part[it] = q < Q ? q % HP4 : 0;
  }
  const uint32_t seg_bytes_site = HP4 * 16;
  const uint32_t stage_bytes = (uint32_t)Q * 16;
  uint32_t phase_bits = 0;                          // parity of each stage's mbarrier (bit i)
  const double k_c = sv.c, k_sigma = sv.y_scale;

  for (;;) {
    long long c = 0;
    if (lane == 0) c = (long long)atomicAdd(P.counters, 1ULL);
    c = __shfl_sync(SO_FULL, c, 0);
    if (c >= P.n_chains) break;
    uint32_t *bits_g = P.bits + c * P.W;
    for (int w = lane; w < P.W; w += 32) s_bits[w] = bits_g[w];
    int4 *hq = (int4 *)P.hq + c * (long long)P.N * HP4;
    long long hi = P.hi[c], lo = P.lo[c];
    const double tsc = P.tscale[c];
    unsigned n_acc = 0, n_guard = 0;
    __syncwarp();

    // lane-parallel draws for 32 steps at a time: batch b covers steps 32b .. 32b+31
    int f_cur = 0, f_nxt = 0;
    float u_cur = 0.f, u_nxt = 0.f;
    auto draw_batch = [&](long long b, int &fo, float &uo) {
      long long s = b * 32 + lane;
      fo = 0; uo = 0.f;
      if (s < P.n_steps) {
        if (P.proposals) {
          fo = __ldg(P.proposals + c * P.n_steps + s);
          uo = __ldg(P.uniforms + c * P.n_steps + s);
        } else {
          unsigned long long step = (unsigned long long)(P.step0 + s), cid = (unsigned long long)(P.chain_id0 + c);
          uint32_t o[4];
          philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), (uint32_t)cid, (uint32_t)(cid >> 32), (uint32_t)P.seed,
                        (uint32_t)(P.seed >> 32), o);
          fo = (int)__umulhi(o[0], (uint32_t)P.N);
          uo = (float)(o[1] >> 8) * 5.9604644775390625e-08f;
        }
      }
    };
    draw_batch(0, f_cur, u_cur);
    draw_batch(1, f_nxt, u_nxt);
    // issue the bulk loads of the window around site f into stage `slot`
    auto issue_loads = [&](int f, int slot) {
      const int f2 = f / d, f1 = f - f2 * d;
      const uint32_t bar = smem_u32(s_bar + slot);
      if (lane == 0) mbar_expect_tx(bar, stage_bytes);
      __syncwarp();
      if (lane < g.n_seg) {
        const int r2 = wrap(f2 - g.b2 + lane, d);
        int start = f1 - g.b1;                                       // first site of the segment, may be < 0
        const uint32_t dst = smem_u32(s_stage + (size_t)slot * stage_i4 + (size_t)lane * g.L * HP4);
        const int4 *row = hq + (size_t)r2 * d * HP4;
        int n1;                                                      // sites before the wrap point
        if (start < 0) { start += d; n1 = min(g.L, d - start); }
        else n1 = min(g.L, d - start);
        bulk_load(dst, row + (size_t)start * HP4, (uint32_t)n1 * seg_bytes_site, bar);
        if (n1 < g.L) bulk_load(dst + (uint32_t)n1 * seg_bytes_site, row, (uint32_t)(g.L - n1) * seg_bytes_site, bar);
      }
    };
    // write the (updated) window in stage `slot` back to global memory
    auto issue_stores = [&](int f, int slot) {
      const int f2 = f / d, f1 = f - f2 * d;
      if (lane < g.n_seg) {
        const int r2 = wrap(f2 - g.b2 + lane, d);
        int start = f1 - g.b1;
        const uint32_t src = smem_u32(s_stage + (size_t)slot * stage_i4 + (size_t)lane * g.L * HP4);
        int4 *row = hq + (size_t)r2 * d * HP4;
        if (start < 0) start += d;
        int n1 = min(g.L, d - start);
        bulk_store(row + (size_t)start * HP4, src, (uint32_t)n1 * seg_bytes_site);
        if (n1 < g.L) bulk_store(row, src + (uint32_t)n1 * seg_bytes_site, (uint32_t)(g.L - n1) * seg_bytes_site);
        bulk_commit();
      }
    };

    // history of flips accepted in the last STAGES-1 steps (for staleness of prefetched windows) and the flip whose
    // stores may still be in flight
    int hist_f[STAGES - 1];
#pragma unroll
    for (int i = 0; i < STAGES - 1; ++i) hist_f[i] = -1;
    int pend_f = -1;
    auto overlaps = [&](int fa, int fb) -> bool {
      if (fa < 0 || fb < 0) return false;
      int a2 = fa / d, a1 = fa - a2 * d, b2 = fb / d, b1 = fb - b2 * d;
      return near_cyclic(a1, b1, g.L - 1, d) && near_cyclic(a2, b2, g.n_seg - 1, d);
    };
    auto proposal = [&](long long s, long long batch0) -> int {   // batch0 = index of the batch held in *_cur
      int v0 = __shfl_sync(SO_FULL, f_cur, (int)(s & 31));
      int v1 = __shfl_sync(SO_FULL, f_nxt, (int)(s & 31));
      return ((s >> 5) == batch0) ? v0 : v1;
    };

    // prologue: windows of steps 0 .. STAGES-2 in flight
    long long batch0 = 0;
#pragma unroll
    for (int i = 0; i < STAGES - 1; ++i)
      if (i < P.n_steps) issue_loads(proposal(i, batch0), i);

    int slot = 0;
    for (long long s = 0; s < P.n_steps; ++s) {
      if ((s >> 5) != batch0) {                      // entered the next batch of draws
        batch0 = s >> 5;
        f_cur = f_nxt; u_cur = u_nxt;
        draw_batch(batch0 + 1, f_nxt, u_nxt);
      }
      const int f = __shfl_sync(SO_FULL, f_cur, (int)(s & 31));
      const float u = __shfl_sync(SO_FULL, u_cur, (int)(s & 31));
      const double T = __dmul_rn(tsc, __ldg(P.temps + s));

      // prefetch step s+STAGES-1 into the stage freed by step s-1
      {
        long long sp = s + STAGES - 1;
        if (sp < P.n_steps) {
          int fp = proposal(sp, batch0);
          int pslot = slot == 0 ? STAGES - 1 : slot - 1;
          if (overlaps(fp, pend_f)) { bulk_wait_all(); pend_f = -1; }   // a store of that region may be in flight
          issue_loads(fp, pslot);
        }
      }
      // is the prefetched window of step s stale?  (a flip accepted after its prefetch was issued overlaps it)
      bool stale = false;
#pragma unroll
      for (int i = 0; i < STAGES - 1; ++i) stale |= overlaps(f, hist_f[i]);
      const uint32_t bar = smem_u32(s_bar + slot);
      mbar_wait(bar, (phase_bits >> slot) & 1u);
      phase_bits ^= (1u << slot);
      if (stale) {
        bulk_wait_all();                             // all stores of this lane complete and visible
        pend_f = -1;
        __syncwarp();
        issue_loads(f, slot);
        mbar_wait(bar, (phase_bits >> slot) & 1u);
        phase_bits ^= (1u << slot);
      }

      const int sgn = ((s_bits[f >> 5] >> (f & 31)) & 1) ? -1 : 1;
      int4 *stg = s_stage + (size_t)slot * stage_i4;
      long long acc = 0;
      int4 hn[MAXIT];
#pragma unroll
      for (int it = 0; it < MAXIT; ++it) {
        int q = lane + 32 * it;
        if (q < Q) {
          int4 h = stg[q];
          int4 w = Wreg[it], w2 = s_w2[part[it]], n;
          n.x = h.x + sgn * w.x; n.y = h.y + sgn * w.y; n.z = h.z + sgn * w.z; n.w = h.w + sgn * w.w;
          acc += (long long)w2.x * (long long)(max(n.x, 0) - max(h.x, 0));
          acc += (long long)w2.y * (long long)(max(n.y, 0) - max(h.y, 0));
          acc += (long long)w2.z * (long long)(max(n.z, 0) - max(h.z, 0));
          acc += (long long)w2.w * (long long)(max(n.w, 0) - max(h.w, 0));
          hn[it] = n;
        }
      }
      // exact warp sum of int64 via three 21-bit limbs (each limb sum fits 32 bits)
      {
        int l0 = (int)(acc & 0x1fffff), l1 = (int)((acc >> 21) & 0x1fffff), l2 = (int)(acc >> 42);
        l0 = __reduce_add_sync(SO_FULL, l0);
        l1 = __reduce_add_sync(SO_FULL, l1);
        l2 = __reduce_add_sync(SO_FULL, l2);
        acc = ((long long)l2 << 42) + ((long long)l1 << 21) + (long long)l0;
      }
      const double delta = so_real(k_c, sv.b2, k_sigma, sv.y_mean, acc >> SO_SPLIT_BITS, acc & SO_SPLIT_MASK, 0);
      // Metropolis (OML/optimize_SA.py:108-116): fp32 pre-filter, fp64 rule when it cannot decide
      bool accept, guard = false;
      if (delta > 0.0) {
        accept = true;
        if (P.exact_guard && T <= 0.0 && delta < P.guard_tol) guard = true;
      } else if (T > 0.0) {
        float pf = __expf((float)delta / (float)T);
        if (!P.exact_guard && u > 0.f && fabsf(pf - u) > 1e-3f * fmaxf(pf, u)) {
          accept = pf > u;
        } else {
          accept = exp(delta / T) > (double)u;
          if (P.exact_guard && u > 0.f) guard = fabs(delta - T * log((double)u)) < P.guard_tol;
        }
      } else {
        accept = false;
        if (P.exact_guard && delta > -P.guard_tol) guard = true;
      }
      if (guard) {
        double de = so_exact_delta(sv, s_bits, d, f, lane);
        if (de > 0.0) accept = true;
        else if (T > 0.0) accept = exp(de / T) > (double)u;
        else accept = false;
        ++n_guard;
      }

      if (accept) {
#pragma unroll
        for (int it = 0; it < MAXIT; ++it)
          if (lane + 32 * it < Q) stg[lane + 32 * it] = hn[it];
        fence_async_smem();                          // generic-proxy writes -> visible to the bulk (async proxy) store
        __syncwarp();
        bulk_wait_all();                             // at most one flip's stores in flight
        issue_stores(f, slot);
        pend_f = f;
        if (lane == 0) s_bits[f >> 5] ^= (1u << (f & 31));
        lo += acc & SO_SPLIT_MASK;
        hi += (acc >> SO_SPLIT_BITS) + (lo >> SO_SPLIT_BITS);
        lo &= SO_SPLIT_MASK;
        ++n_acc;
      }
#pragma unroll
      for (int i = STAGES - 2; i > 0; --i) hist_f[i] = hist_f[i - 1];
      hist_f[0] = accept ? f : -1;
      if (P.accept_out && lane == 0) P.accept_out[c * P.n_steps + s] = accept ? 1 : 0;
      // the stage is re-used by the prefetch issued at the next step: its stores must have read it by then
      if (accept) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      slot = slot + 1 == STAGES ? 0 : slot + 1;
    }
    bulk_wait_all();
    __syncwarp();
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

}  // namespace

size_t so_window_smem_bytes(int Q, int W, int stages) {
  size_t i4 = 8 + (size_t)WARPS * stages * Q;
  size_t words = ((size_t)WARPS * W + 3) & ~(size_t)3;
  return i4 * 16 + words * 4 + (size_t)WARPS * stages * 8 + 16;
}

int so_launch_anneal_window(const so_chains *c, const SaParams &P, cudaStream_t st) {
  const so_surrogate *s = c->sur;
  int n_sm = 0;
  SO_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, c->device));
  WinGeom g;
  g.b1 = s->win_b1; g.b2 = s->win_b2; g.L = s->win_b1 - s->win_a1 + 1; g.n_seg = s->win_b2 - s->win_a2 + 1;
  const int Q = s->K * s->HP4;
  constexpr int STAGES = 3;
  size_t smem = so_window_smem_bytes(Q, P.W, STAGES);
  long long want = (P.n_chains + WARPS - 1) / WARPS;
  int blocks = (int)(want < (long long)n_sm * 2 ? want : (long long)n_sm * 2);
  if (blocks < 1) blocks = 1;
#define SO_WIN_CASE(HP4V)                                                                                        \
  case HP4V: {                                                                                                   \
    auto kern = sa_window_kernel<HP4V, STAGES>;                                                                  \
    SO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
    kern<<<blocks, WARPS * 32, smem, st>>>(P, g, (const int4 *)s->Wst_dev);                                      \
  } break;
  switch (s->HP4) {
    SO_WIN_CASE(1) SO_WIN_CASE(2) SO_WIN_CASE(3) SO_WIN_CASE(4) SO_WIN_CASE(5) SO_WIN_CASE(6) SO_WIN_CASE(7) SO_WIN_CASE(8)
    default: so_set_error("hidden width not supported"); return SO_EINVAL;
  }
#undef SO_WIN_CASE
  SO_CUDA(cudaGetLastError());
  return SO_OK;
}
