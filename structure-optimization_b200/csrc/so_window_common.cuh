// so_window_common.cuh -- building blocks of the pipelined window kernel (so_sa_window.cu): mbarrier / bulk-async /
// tensor-TMA wrappers, the window geometry, and the out-of-line cold paths.
#pragma once
#include <cuda.h>
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
  int use_tma;       // interior windows go through one tensor-TMA box
  int stage_i4;      // int4 per shared-memory stage: window + zero-weight padding, multiple of 8 (128 B)
  int nslots;        // lane slots per window: ceil(Q / STRIDE)
  int tab_i4;        // gated kernel: entries of the weight tables (state slots + zero padding)
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap *tmap, int c0, int c1, int c2,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst_smem), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *tmap, int c0, int c1, int c2, uint32_t src_smem) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap),
               "r"(src_smem), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// packed site coordinates: f1 | f2 << 16
__device__ __forceinline__ int pk1(int p) { return p & 0xffff; }
__device__ __forceinline__ int pk2(int p) { return p >> 16; }
__device__ __forceinline__ bool near_cyclic(int a, int b, int r, int d) {   // cyclic |a - b| <= r
  int t = abs(a - b);
  return min(t, d - t) <= r;
}

// ---- cold paths, kept out of line so that they do not cost the hot loop registers ---------------------------------
// lane-parallel draws of steps 32b .. 32b+31: packed site (f1 | f2 << 16) and the fp32 DECISION THRESHOLD of the step.
// The Metropolis rule (OML/optimize_SA.py:108-116) "accept iff delta > 0, or T > 0 and exp(delta / T) > u" is
// "accept iff delta > thr" with thr = min(T ln u, 0) for T > 0 and u > 0, thr = 0 for T <= 0 (x = 0 is then a reject and is
// left to the fp64 rule).  thr is computed here in fp64, once per step and lane, and rounded to fp32; the hot loop compares
// the fp32 image of delta with it and falls back to the fp64 rule (decide_cold) whenever the two are closer than 1e-5
// relative -- 50x the worst fp32 error of either side -- so every decision is the fp64 rule's.  thr = NaN (u = 0, exact_guard,
// an fp32-denormal scale) sends the step to the fp64 rule unconditionally.
// Everything is passed and returned BY VALUE: an address-taken argument would pin the caller's variables in local memory.
struct Draw {
  int f;
  float thr;
};
__device__ __forceinline__ float draw_u(const float *uniforms, long long n_steps, long long step0, long long chain_id,
                                        long long c, unsigned long long seed, long long s, int N, int *f_out) {
  if (uniforms) return __ldg(uniforms + c * n_steps + s);
  unsigned long long step = (unsigned long long)(step0 + s), cid = (unsigned long long)chain_id;
  uint32_t o[4];
  philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), (uint32_t)cid, (uint32_t)(cid >> 32), (uint32_t)seed,
                (uint32_t)(seed >> 32), o);
  if (f_out) *f_out = (int)__umulhi(o[0], (uint32_t)N);
  return (float)(o[1] >> 8) * 5.9604644775390625e-08f;
}
__device__ __noinline__ Draw draw_batch_cold(const double *temps, const int32_t *proposals, const float *uniforms,
                                             long long n_steps, long long step0, long long chain_id, long long c,
                                             unsigned long long seed, int N, int d, long long b, int lane, double tsc,
                                             int fast_ok) {
  long long s = b * 32 + lane;
  Draw r;
  int f = 0;
  r.thr = __int_as_float(0x7fc00000);
  if (s < n_steps) {
    const float u = draw_u(uniforms, n_steps, step0, chain_id, c, seed, s, N, &f);
    if (proposals) f = __ldg(proposals + c * n_steps + s);
    const double T = __dmul_rn(tsc, __ldg(temps + s));
    if (fast_ok) {
      if (T > 0.0) {
        if (u > 0.f) r.thr = (float)fmin(T * log((double)u), 0.0);
      } else {
        r.thr = 0.f;
      }
    }
  }
  int f2 = f / d;
  r.f = (f - f2 * d) | (f2 << 16);
  return r;
}
// the uniform of ONE step again, for the fp64 rule (the hot loop carries only the threshold)
__device__ __noinline__ float redraw_u_cold(const float *uniforms, long long n_steps, long long step0, long long chain_id,
                                            long long c, unsigned long long seed, long long s) {
  return draw_u(uniforms, n_steps, step0, chain_id, c, seed, s, 1, nullptr);
}

// the fp64 Metropolis rule (OML/optimize_SA.py:108-116) + the near-tie guard; bit 0 = accept, bit 1 = guard
__device__ __noinline__ int metropolis_cold(double delta, double T, float u, int exact_guard, double tol) {
  bool accept;
  if (delta > 0.0) accept = true;
  else if (T > 0.0) accept = exp(delta / T) > (double)u;
  else accept = false;
  const bool guard = exact_guard && so_near_tie(delta, T, u, tol);
  return (accept ? 1 : 0) | (guard ? 2 : 0);
}

__device__ __noinline__ int exact_decision_cold(const SurView sv, const uint32_t *bits, int d, int f, int lane, double T,
                                                float u) {
  double de = so_exact_delta(sv, bits, d, f, lane);
  if (de > 0.0) return 1;
  if (T > 0.0) return exp(de / T) > (double)u ? 1 : 0;
  return 0;
}

// windows that cross the periodic boundary (about 2*(L+n_seg)/d of all flips): plain bulk copies, one window row per
// LANE (lanes 0 .. n_seg-1 each issue the one or two copies of their row; all complete on the same mbarrier, whose
// expect_tx lane 0 posts for the whole window) -- a single pass of the warp instead of a loop in one lane
__device__ __noinline__ void boundary_loads_cold(const int4 *hq, uint32_t dst, uint32_t bar, int f1, int f2, int b1,
                                                 int b2, int L, int n_seg, int d, int hp4, int lane) {
  const uint32_t site_bytes = hp4 * 16;
  int start = f1 - b1;
  if (start < 0) start += d;
  const int n1 = min(L, d - start);
  const int sg = lane;
  if (sg < n_seg) {
    const int4 *row = hq + (size_t)wrap(f2 - b2 + sg, d) * d * hp4;
    const uint32_t dseg = dst + (uint32_t)(sg * L) * site_bytes;
    bulk_load(dseg, row + (size_t)start * hp4, (uint32_t)n1 * site_bytes, bar);
    if (n1 < L) bulk_load(dseg + (uint32_t)n1 * site_bytes, row, (uint32_t)(L - n1) * site_bytes, bar);
  }
}
// the same for the write-back of an accepted flip: lanes 0 .. n_seg-1 each store their row; every issuing lane owns its
// own bulk group (commit / wait are per thread), so ALL lanes call this and then commit + wait themselves
__device__ __noinline__ void boundary_stores_cold(int4 *hq, uint32_t src, int f1, int f2, int b1, int b2, int L, int n_seg,
                                                  int d, int hp4, int lane) {
  const uint32_t site_bytes = hp4 * 16;
  int start = f1 - b1;
  if (start < 0) start += d;
  const int n1 = min(L, d - start);
  const int sg = lane;
  if (sg < n_seg) {
    int4 *row = hq + (size_t)wrap(f2 - b2 + sg, d) * d * hp4;
    const uint32_t sseg = src + (uint32_t)(sg * L) * site_bytes;
    bulk_store(row + (size_t)start * hp4, sseg, (uint32_t)n1 * site_bytes);
    if (n1 < L) bulk_store(row, sseg + (uint32_t)n1 * site_bytes, (uint32_t)(L - n1) * site_bytes);
  }
}
// a prefetched window went stale (a flip accepted after its prefetch overlaps it): all lanes re-read it from L2
// with plain loads once the pending stores have drained
__device__ __noinline__ void reload_window_cold(int4 *stage, const int4 *hq, int f1, int f2, int b1, int b2, int L, int Q,
                                                int d, int hp4, int lane) {
  for (int q = lane; q < Q; q += 32) {
    int p = q / hp4, part = q - p * hp4;
    int p2 = p / L, p1 = p - p2 * L;
    int r1 = wrap(f1 - b1 + p1, d), r2 = wrap(f2 - b2 + p2, d);
    stage[q] = __ldcg(hq + ((size_t)r2 * d + r1) * hp4 + part);
  }
}


}  // namespace
