// so_sa_resident.cu -- SA for SMALL lattices (the reference's own regime, d = 12 / 16): the whole first-layer state of a
// chain (4*N*Hp bytes, 11.5 KB at N = 144, H = 20) lives in shared memory for the entire launch, so a step touches no
// HBM at all.  Same arithmetic and decisions as the other SA kernels (exact integer sums, fp64 Metropolis rule), any
// descriptor (full translation row or window), with or without the tree gate.
//
//   * one warp per chain; the CTA shares the weight table W1q, the descriptor offsets and the tree;
//   * delta of the rows that are active before AND after the flip: lanes own a fixed hidden-unit quad and stride over
//     the affected rows (32/HP4 rows per pass), accumulating relu differences per component in int32 (a column's
//     differences are bounded by sum_k |W1q_kj| < 2^30); the second-layer weights enter once per step, the warp sum is
//     three REDUX on 21-bit limbs;
//   * tree gate: two caches per chain make a step almost walk-free.  (1) the gate of every row of the CURRENT
//     structure, one bit per row.  (2) the path index S[v] (K bits per site v): bit c is set iff the current path of
//     row v - offset_c tests column c, i.e. site v.  A flip at f can only change the gates of the rows in S[f] (about
//     depth/K of them), so only those walk the tree; on accept their old path entries are dropped and the new ones
//     entered.  Rows whose gate changes contribute their whole row sum (rare, out of line);
//   * draws come lane-parallel 32 steps at a time (Philox or the injected sequence); the decision uses the fp32
//     pre-filter of the window kernel and the fp64 rule when that cannot decide.
#include "so_internal.cuh"
#include "so_types.cuh"
#include <math.h>
#include <stdlib.h>

#define SO_TREE_SMEM_NODES 2048

namespace {

struct DrawR {
  int f;        // packed (f1 | f2 << 16)
  float u;
  float tf;
};
// lane-parallel draws of steps 32b .. 32b+31; everything by value (an address-taken argument would pin the caller's
// variables in local memory)
__device__ __noinline__ DrawR draw_batch_res(const double *temps, const int32_t *proposals, const float *uniforms,
                                             long long n_steps, long long step0, long long chain_id, long long c,
                                             unsigned long long seed, int N, int d, long long b, int lane) {
  long long s = b * 32 + lane;
  DrawR r;
  int f = 0;
  r.u = 0.f;
  r.tf = 0.f;
  if (s < n_steps) {
    r.tf = (float)__ldg(temps + s);
    if (proposals) {
      f = __ldg(proposals + c * n_steps + s);
      r.u = __ldg(uniforms + c * n_steps + s);
    } else {
      unsigned long long step = (unsigned long long)(step0 + s), cid = (unsigned long long)chain_id;
      uint32_t o[4];
      philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), (uint32_t)cid, (uint32_t)(cid >> 32), (uint32_t)seed,
                    (uint32_t)(seed >> 32), o);
      f = (int)__umulhi(o[0], (uint32_t)N);
      r.u = (float)(o[1] >> 8) * 5.9604644775390625e-08f;
    }
  }
  int f2 = f / d;
  r.f = (f - f2 * d) | (f2 << 16);
  return r;
}

// the fp64 Metropolis rule (OML/optimize_SA.py:108-116) + near-tie guard test; bit 0 accept, bit 1 guard
__device__ __noinline__ int decide_res(double delta, double tsc, const double *temps, long long s, float u, int exact_guard,
                                       double tol) {
  const double T = __dmul_rn(tsc, __ldg(temps + s));
  bool accept;
  if (delta > 0.0) accept = true;
  else if (T > 0.0) accept = exp(delta / T) > (double)u;
  else accept = false;
  const bool guard = exact_guard && so_near_tie(delta, T, u, tol);
  return (accept ? 1 : 0) | (guard ? 2 : 0);
}
__device__ __noinline__ int exact_decision_res(const SurView sv, const uint32_t *bits, int d, int f, int lane, double tsc,
                                               const double *temps, long long s, float u) {
  const double T = __dmul_rn(tsc, __ldg(temps + s));
  double de = so_exact_delta(sv, bits, d, f, lane);
  if (de > 0.0) return 1;
  if (T > 0.0) return exp(de / T) > (double)u ? 1 : 0;
  return 0;
}

// whole row sum  sum_j w2q_j * relu(hq_rj + sgn * W1q_kj)  (sgn = 0: the row as it is); rows whose gate changes only
template <int HP4>
__device__ __forceinline__ long long row_sum_res(const int4 *hrow, const int4 *wrow, const int4 *w2s, int sgn) {
  long long A = 0, B = 0;
#pragma unroll
  for (int p = 0; p < HP4; ++p) {
    const int4 h = hrow[p], w = wrow[p], w2 = w2s[p];
    A += (long long)w2.x * max(h.x + sgn * w.x, 0) + (long long)w2.y * max(h.y + sgn * w.y, 0);
    B += (long long)w2.z * max(h.z + sgn * w.z, 0) + (long long)w2.w * max(h.w + sgn * w.w, 0);
  }
  return A + B;
}

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__host__ __device__ __forceinline__ size_t a16(size_t v) { return (v + 15) & ~(size_t)15; }

// Walk row (r1, r2) through the tree for the structure `bits` with site f read inverted (f = -1: as it is).
__device__ __forceinline__ int walk1(const int4 *tree, const uint32_t *bits, int d, int f, int r1, int r2) {
  int4 nd = tree[0];
  while (!(nd.w & 1)) {
    const int o1 = (int)(short)(nd.x & 0xffff), o2 = nd.x >> 16;
    const int v = wrap(r2 + o2, d) * d + wrap(r1 + o1, d);
    const int bit = (int)((bits[v >> 5] >> (v & 31)) & 1) ^ (v == f);
    nd = tree[((nd.w >> (2 + bit)) & 1) ? nd.y : nd.z];
  }
  return (nd.w >> 1) & 1;
}
// The same walk on the structure as it is, entering (SET) or dropping the path in the index S: every node whose test
// depends on its bit marks (site tested, column tested).  Returns the gate.
template <bool SET>
__device__ __noinline__ int walk_mark(const int4 *tree, const uint32_t *bits, uint32_t *S, int KW, int d, int r1, int r2) {
  int4 nd = tree[0];
  while (!(nd.w & 1)) {
    const int o1 = (int)(short)(nd.x & 0xffff), o2 = nd.x >> 16;
    const int v = wrap(r2 + o2, d) * d + wrap(r1 + o1, d);
    const int bit = (int)((bits[v >> 5] >> (v & 31)) & 1);
    if (((nd.w >> 2) ^ (nd.w >> 3)) & 1) {
      const int c = nd.w >> 8;
      if (SET) atomicOr(S + v * KW + (c >> 5), 1u << (c & 31));
      else atomicAnd(S + v * KW + (c >> 5), ~(1u << (c & 31)));
    }
    nd = tree[((nd.w >> (2 + bit)) & 1) ? nd.y : nd.z];
  }
  return (nd.w >> 1) & 1;
}

// RES: first-layer state (and the path index) of the chain in shared memory for the whole launch.  !RES (large
// lattices, gated only): the same step with hq in global memory -- only the rows that are active are ever read --
// and the path index in a per-warp global scratch; occupancy and gate bits stay in shared memory.
template <int HP4, bool GATED, bool RES>
__global__ void __launch_bounds__(RES ? 512 : 256, RES ? 1 : 4) sa_resident_kernel(SaParams P) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  constexpr int RPP = 32 / HP4;          // rows per pass
  constexpr int STRIDE = RPP * HP4;      // working lanes
  const SurView &sv = P.sv;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = P.d, N = P.N, K = sv.K, KW = (K + 31) >> 5, W = P.W;
  // trees of up to SO_TREE_SMEM_NODES nodes are copied to shared memory, larger ones are walked in global memory (L1/L2)
  const int n_tree = (GATED && sv.n_tree <= SO_TREE_SMEM_NODES) ? sv.n_tree : 0;
  // shared layout: [W1q: K*HP4 int4][tree: n_tree int4][offsets: K][w2q: HP4 int4] | per warp: [hq: N*HP4 int4][bits W][act W]
  // [path index S: N*KW][active rows of the step: K u32][their positions: K u16]
  int4 *s_W = (int4 *)s_raw;
  int4 *s_tree_buf = s_W + (size_t)K * HP4;
  const int4 *s_tree = n_tree ? s_tree_buf : sv.t_packed;
  int32_t *s_off = (int32_t *)(s_tree_buf + n_tree);
  int4 *s_w2 = (int4 *)((unsigned char *)s_off + a16((size_t)K * 4));
  unsigned char *s_chain0 = (unsigned char *)(s_w2 + HP4);
  // resident state with the path index in the per-warp GLOBAL scratch (variant_flags bit 3): 2.9 KB less shared memory per warp
  // at N = K = 144 buys 16 instead of 12 warps per SM for a latency-bound step
  const bool s_glob = GATED && RES && (P.variant_flags & 8);
  const bool s_smem = GATED && RES && !s_glob;
  const size_t per_warp = (RES ? (size_t)N * HP4 * 16 : 0) +
                          a16((size_t)(2 * W + (GATED ? (s_smem ? N * KW : 0) + K : 0)) * 4 + (GATED ? (size_t)K * 2 : 0));
  int4 *s_hq = (int4 *)(s_chain0 + (size_t)warp * per_warp);
  uint32_t *s_bits = (uint32_t *)(s_hq + (RES ? (size_t)N * HP4 : 0));
  uint32_t *s_act = s_bits + W;
  uint32_t *s_S;
  if (s_smem) s_S = s_act + W;
  else s_S = P.gate_scratch + (size_t)(blockIdx.x * (blockDim.x >> 5) + warp) * N * KW;
  uint32_t *s_list = s_act + W + (s_smem ? N * KW : 0);         // compacted active rows of the step (gated)
  unsigned short *s_pos = (unsigned short *)(s_list + (GATED ? K : 0));
  for (int q = threadIdx.x; q < K * HP4; q += blockDim.x) s_W[q] = ((const int4 *)sv.W1q)[q];
  for (int k = threadIdx.x; k < K; k += blockDim.x) s_off[k] = sv.offs[k];
  if (threadIdx.x < HP4) s_w2[threadIdx.x] = ((const int4 *)sv.w2q)[threadIdx.x];
  if (GATED) {
    for (int i = threadIdx.x; i < n_tree; i += blockDim.x) s_tree_buf[i] = sv.t_packed[i];
  }
  __syncthreads();
  const int part = lane % HP4, rsub = lane / HP4;
  const bool worker = lane < STRIDE;
  const int4 w2 = worker ? __ldg((const int4 *)sv.w2q + part) : make_int4(0, 0, 0, 0);
  const float tol_f = (float)P.guard_tol;
  const bool guard_on = P.exact_guard != 0;
  // FULL translation descriptor (K = N, column k <-> offset (k % d, k / d)): every row is affected by every flip, so the
  // rows that are active are kept in a PERSISTENT compact list (s_list: row | r1 << 16 | r2 << 24, s_pos: position of a
  // row) that only changes when an accepted flip switches a gate -- the per-step pass over the K rows' gates disappears
  const bool full = GATED && RES && (P.variant_flags & 2);

  for (;;) {
    long long c = 0;
    if (lane == 0) c = (long long)atomicAdd(P.counters, 1ULL);
    c = __shfl_sync(SO_FULL, c, 0);
    if (c >= P.n_chains) break;
    uint32_t *bits_g = P.bits + c * W;
    int4 *hq_g = (int4 *)P.hq + c * (long long)N * HP4;
    for (int w = lane; w < W; w += 32) s_bits[w] = bits_g[w];
    int4 *hq;
    if constexpr (RES) {
      for (int q = lane; q < N * HP4; q += 32) s_hq[q] = hq_g[q];
      hq = s_hq;
    } else {
      hq = hq_g;
    }
    long long hi = P.hi[c], lo = P.lo[c];
    int nact = P.nact[c];
    const double tsc = P.tscale[c];
    const float tsc_f = (float)tsc;
    unsigned n_acc = 0, n_guard = 0;
    int n_alist = 0;               // FULL descriptor: number of active rows in the persistent list
    __syncwarp();
    if (GATED) {      // gates and path index of the starting structure: row r = 32 j + lane -> bit `lane` of word j
      for (int i = lane; i < N * KW; i += 32) s_S[i] = 0u;
      __syncwarp();
      for (int r0 = 0; r0 < N; r0 += 32) {
        const int r = r0 + lane;
        int g = 0;
        if (r < N) {
          const int r2 = r / d;
          g = walk_mark<true>(s_tree, s_bits, s_S, KW, d, r - r2 * d, r2);
        }
        __syncwarp();
        const unsigned m = __ballot_sync(SO_FULL, g);
        if (lane == 0) s_act[r0 >> 5] = m;
        if (full) {
          if (g) {
            const int pos = n_alist + __popc(m & ((1u << lane) - 1u));
            const int r2 = r / d;
            s_list[pos] = (unsigned)r | ((unsigned)(r - r2 * d) << 16) | ((unsigned)r2 << 24);
            s_pos[r] = (unsigned short)pos;
          }
          n_alist += __popc(m);
        }
      }
      __syncwarp();
    }
    DrawR batch;
    batch.f = 0; batch.u = 0.f; batch.tf = 0.f;

    for (long long s = 0; s < P.n_steps; ++s) {
      if ((s & 31) == 0)
        batch = draw_batch_res(P.temps, P.proposals, P.uniforms, P.n_steps, P.step0, P.chain_id0 + c, c, P.seed, N, d,
                               s >> 5, lane);
      const int fp = __shfl_sync(SO_FULL, batch.f, (int)(s & 31));
      const float u = __shfl_sync(SO_FULL, batch.u, (int)(s & 31));
      const float Tf = tsc_f * __shfl_sync(SO_FULL, batch.tf, (int)(s & 31));
      const int f1 = fp & 0xffff, f2 = fp >> 16;
      const int f = f2 * d + f1;
      const int sgn = ((s_bits[f >> 5] >> (f & 31)) & 1) ? -1 : 1;
      if (s_glob && ((s + 1) & 31) != 0) {         // index in global memory: the next step's words are pulled towards the SM now
        const int fn = __shfl_sync(SO_FULL, batch.f, (int)((s + 1) & 31));
        if (lane < KW) prefetch_l1(s_S + ((fn >> 16) * d + (fn & 0xffff)) * KW + lane);
      }
      if constexpr (!RES) {
        // global state: pull the next step's index words and active rows towards the SM while this step computes
        // (a hint only; the draws of a batch are known 32 steps ahead)
        if (((s + 1) & 31) != 0) {
          const int fn = __shfl_sync(SO_FULL, batch.f, (int)((s + 1) & 31));
          const int n1 = fn & 0xffff, n2 = fn >> 16;
          if (lane < KW) prefetch_l1(s_S + (n2 * d + n1) * KW + lane);
          for (int j = 0; j < KW; ++j) {
            const int k = 32 * j + lane;
            if (k < K) {
              const int off = s_off[k];
              const int row = wrap(n2 - (off >> 16), d) * d + wrap(n1 - (int)(short)(off & 0xffff), d);
              if ((s_act[row >> 5] >> (row & 31)) & 1u) {
                const char *p = (const char *)(hq + row * HP4);
                prefetch_l1(p);
                if ((((size_t)p) & 127) + HP4 * 16 > 128) prefetch_l1(p + HP4 * 16 - 1);
              }
            }
          }
        }
      }
      long long dhi = 0;
      int dlo = 0, dn = 0;

      // per-lane bit masks over the lane's rows (row of group j: k = 32 j + lane): path tests f / gate changes
      unsigned mywalk = 0, mychg = 0, myon = 0;
      int n_list = 0;            // rows active before the flip, compacted: s_list[i] = k << 16 | live << 15 | row
      if (GATED && full) {
        n_list = n_alist;
        for (int j = 0; j < KW; ++j) mywalk |= ((s_S[f * KW + j] >> lane) & 1u) << j;
        // only rows whose current path tests site f can change their gate: walk those for the flipped structure
        for (unsigned todo = mywalk; todo; todo &= todo - 1) {
          const int j = __ffs(todo) - 1;
          const int off = s_off[32 * j + lane];
          const int r1 = wrap(f1 - (int)(short)(off & 0xffff), d), r2 = wrap(f2 - (off >> 16), d);
          const int row = r2 * d + r1;
          const unsigned g0 = (s_act[row >> 5] >> (row & 31)) & 1u;
          const unsigned g1 = walk1(s_tree, s_bits, d, f, r1, r2);
          mychg |= (g0 ^ g1) << j;
          myon |= g1 << j;
        }
        __syncwarp();
        // a row that switches contributes its whole NEW row sum: + if it switches on (it is not in the list), - if it
        // switches off (it stays in the list for this step and contributes new - old there: the total is - old)
        for (unsigned todo = mychg; todo; todo &= todo - 1) {
          const int j = __ffs(todo) - 1, k = 32 * j + lane;
          const int off = s_off[k];
          const int row = wrap(f2 - (off >> 16), d) * d + wrap(f1 - (int)(short)(off & 0xffff), d);
          const long long A = row_sum_res<HP4>(hq + row * HP4, s_W + k * HP4, s_w2, sgn);
          const int sg = ((myon >> j) & 1) ? 1 : -1;
          dhi += sg * (A >> SO_SPLIT_BITS);
          dlo += sg * (int)(A & SO_SPLIT_MASK);
          dn += sg;
        }
      } else if (GATED) {
        unsigned g0m = 0;
        const unsigned lt = (1u << lane) - 1u;
#pragma unroll 2
        for (int j = 0; j < KW; ++j) {
          const int k = 32 * j + lane;
          const bool valid = k < K;
          const int off = s_off[valid ? k : 0];
          const int row = wrap(f2 - (off >> 16), d) * d + wrap(f1 - (int)(short)(off & 0xffff), d);
          const unsigned g0 = valid ? (s_act[row >> 5] >> (row & 31)) & 1u : 0u;
          g0m |= g0 << j;
          mywalk |= ((s_S[f * KW + j] >> lane) & 1u) << j;
          const unsigned m = __ballot_sync(SO_FULL, g0);
          if (g0) {
            const int pos = n_list + __popc(m & lt);
            s_list[pos] = (unsigned)(k << 16) | 0x8000u | (unsigned)row;
            s_pos[k] = (unsigned short)pos;
          }
          n_list += __popc(m);
        }
        // only rows whose current path tests site f can change their gate: walk those for the flipped structure
        unsigned g1m = g0m;
        for (unsigned todo = mywalk; todo; todo &= todo - 1) {
          const int j = __ffs(todo) - 1;
          const int off = s_off[32 * j + lane];
          const unsigned g1 = walk1(s_tree, s_bits, d, f, wrap(f1 - (int)(short)(off & 0xffff), d), wrap(f2 - (off >> 16), d));
          g1m = (g1m & ~(1u << j)) | (g1 << j);
        }
        __syncwarp();
        mychg = g0m ^ g1m;
        // rows that switch on or off leave the difference pass and contribute their whole row sum
        for (unsigned todo = mychg; todo; todo &= todo - 1) {
          const int j = __ffs(todo) - 1, k = 32 * j + lane;
          const int off = s_off[k];
          const int row = wrap(f2 - (off >> 16), d) * d + wrap(f1 - (int)(short)(off & 0xffff), d);
          const int on = (g1m >> j) & 1;
          if (!on) s_list[s_pos[k]] &= ~0x8000u;
          const long long A = row_sum_res<HP4>(hq + row * HP4, s_W + k * HP4, s_w2, on ? sgn : 0);
          const int sg = on ? 1 : -1;
          dhi += sg * (A >> SO_SPLIT_BITS);
          dlo += sg * (int)(A & SO_SPLIT_MASK);
          dn += sg;
        }
        __syncwarp();
      }

      // relu differences of the rows active on both sides (all rows without a gate)
      long long acc;
      if (RES && (P.variant_flags & 1)) {
        // state in shared memory: a lane takes WHOLE rows (HP4 x LDS.128 of the state + of the weights, 12 integer ops
        // per quad) and keeps one int32 accumulator per hidden unit; the second-layer weights enter once per step.
        // (5 lanes sharing a row, as below, spend most of their instructions on the per-row bookkeeping.)
        int D[HP4 * 4];
#pragma unroll
        for (int j = 0; j < HP4 * 4; ++j) D[j] = 0;
        if (GATED) {
          for (int i = lane; i < n_list; i += 32) {
            const unsigned e = s_list[i];
            int row, k, sg;
            if (full) {
              row = e & 0xffffu; sg = sgn;
              k = wrap(f2 - (int)(e >> 24), d) * d + wrap(f1 - (int)((e >> 16) & 0xffu), d);
            } else {
              row = e & 0x7fffu; k = e >> 16;
              sg = sgn & -(int)((e >> 15) & 1u);              // branch-free: a row that just switched off sees weight 0
            }
            const int4 *hr = hq + row * HP4, *wr = s_W + k * HP4;
#pragma unroll
            for (int p = 0; p < HP4; ++p) {
              const int4 h = hr[p], w = wr[p];
              D[4 * p + 0] += max(h.x + sg * w.x, 0) - max(h.x, 0);
              D[4 * p + 1] += max(h.y + sg * w.y, 0) - max(h.y, 0);
              D[4 * p + 2] += max(h.z + sg * w.z, 0) - max(h.z, 0);
              D[4 * p + 3] += max(h.w + sg * w.w, 0) - max(h.w, 0);
            }
          }
        } else {
          for (int k = lane; k < K; k += 32) {
            const int off = s_off[k];
            const int row = wrap(f2 - (off >> 16), d) * d + wrap(f1 - (int)(short)(off & 0xffff), d);
            const int4 *hr = hq + row * HP4, *wr = s_W + k * HP4;
#pragma unroll
            for (int p = 0; p < HP4; ++p) {
              const int4 h = hr[p], w = wr[p];
              D[4 * p + 0] += max(h.x + sgn * w.x, 0) - max(h.x, 0);
              D[4 * p + 1] += max(h.y + sgn * w.y, 0) - max(h.y, 0);
              D[4 * p + 2] += max(h.z + sgn * w.z, 0) - max(h.z, 0);
              D[4 * p + 3] += max(h.w + sgn * w.w, 0) - max(h.w, 0);
            }
          }
        }
        acc = 0;
#pragma unroll
        for (int p = 0; p < HP4; ++p) {
          const int4 v = s_w2[p];
          acc += (long long)v.x * D[4 * p + 0] + (long long)v.y * D[4 * p + 1] + (long long)v.z * D[4 * p + 2] +
                 (long long)v.w * D[4 * p + 3];
        }
      } else {
        int ax = 0, ay = 0, az = 0, aw = 0;
        if (worker) {
          if (GATED) {
#pragma unroll 4
            for (int i = rsub; i < n_list; i += RPP) {
              const unsigned e = s_list[i];
              int row, k, sg;
              if (full) {
                row = e & 0xffffu; sg = sgn;
                k = wrap(f2 - (int)(e >> 24), d) * d + wrap(f1 - (int)((e >> 16) & 0xffu), d);
              } else {
                row = e & 0x7fffu; k = e >> 16;
                sg = sgn & -(int)((e >> 15) & 1u);            // branch-free: a row that just switched off sees weight 0
              }
              const int4 h = hq[row * HP4 + part], w = s_W[k * HP4 + part];
              ax += max(h.x + sg * w.x, 0) - max(h.x, 0);
              ay += max(h.y + sg * w.y, 0) - max(h.y, 0);
              az += max(h.z + sg * w.z, 0) - max(h.z, 0);
              aw += max(h.w + sg * w.w, 0) - max(h.w, 0);
            }
          } else {
#pragma unroll 4
            for (int k = rsub; k < K; k += RPP) {
              const int off = s_off[k];
              const int row = wrap(f2 - (off >> 16), d) * d + wrap(f1 - (int)(short)(off & 0xffff), d);
              const int4 h = hq[row * HP4 + part], w = s_W[k * HP4 + part];
              ax += max(h.x + sgn * w.x, 0) - max(h.x, 0);
              ay += max(h.y + sgn * w.y, 0) - max(h.y, 0);
              az += max(h.z + sgn * w.z, 0) - max(h.z, 0);
              aw += max(h.w + sgn * w.w, 0) - max(h.w, 0);
            }
          }
        }
        acc = (long long)w2.x * ax + (long long)w2.y * ay + (long long)w2.z * az + (long long)w2.w * aw;
      }
      {  // exact warp sum: |acc| < 2^62 -> three 21-bit limbs, one REDUX each
        int l0 = (int)(acc & 0x1fffff), l1 = (int)((acc >> 21) & 0x1fffff), l2 = (int)(acc >> 42);
        l0 = __reduce_add_sync(SO_FULL, l0);
        l1 = __reduce_add_sync(SO_FULL, l1);
        l2 = __reduce_add_sync(SO_FULL, l2);
        acc = ((long long)l2 << 42) + ((long long)l1 << 21) + (long long)l0;
      }
      long long dlo_w = acc & SO_SPLIT_MASK;
      if (GATED && __any_sync(SO_FULL, mychg != 0)) {
        int l0 = (int)(dhi & 0x1fffff), l1 = (int)((dhi >> 21) & 0x1fffff), l2 = (int)(dhi >> 42);
        l0 = __reduce_add_sync(SO_FULL, l0);
        l1 = __reduce_add_sync(SO_FULL, l1);
        l2 = __reduce_add_sync(SO_FULL, l2);
        dhi = ((long long)l2 << 42) + ((long long)l1 << 21) + (long long)l0;
        dn = __reduce_add_sync(SO_FULL, dn);
        // |dlo| per lane < 2^20 * rows per lane: the warp sum fits int32 for K <= 1024
        dlo_w += (long long)__reduce_add_sync(SO_FULL, dlo);
      }
      dhi += acc >> SO_SPLIT_BITS;
      const double delta = so_real(sv.c, sv.b2, sv.y_scale, sv.y_mean, dhi, dlo_w, dn);
      bool accept;
      {
        // fp32 pre-filter: decides when p = exp(delta/T) is at least 1e-3 (relative) away from u; with the exact guard
        // on, only where that margin also rules out |delta - T ln u| < tol  (the margin means |ln p - ln u| > 9.8e-4, so tol / T < 5e-4 is safe)
        const float pf = __expf(__fdividef((float)delta, Tf));
        if (delta > 0.0 && (!guard_on || delta >= P.guard_tol)) {     // (with the guard: only when no tie is possible)
          accept = true;
        } else if (delta <= 0.0 && Tf > 1e-30f && u > 0.f && fabsf(pf - u) > 1e-3f * fmaxf(pf, u) &&
                   (!guard_on || tol_f < 5e-4f * Tf)) {
          accept = pf > u;
        } else {
          int r = decide_res(delta, tsc, P.temps, s, u, P.exact_guard, P.guard_tol);
          if (r & 2) {
            r = exact_decision_res(sv, s_bits, d, f, lane, tsc, P.temps, s, u);
            ++n_guard;
          }
          accept = r & 1;
        }
      }
      if (accept) {
        if (P.type_hist) so_track_flip(P.type_hist, P.n_chains, c, s_bits, d, N, f, lane);        // before the bit flips
        if (GATED && __any_sync(SO_FULL, mywalk != 0)) {     // paths through f are about to change: drop them from the index
          for (int j = 0; j < KW; ++j)
            if ((mywalk >> j) & 1) {
              const int off = s_off[32 * j + lane];
              walk_mark<false>(s_tree, s_bits, s_S, KW, d, wrap(f1 - (int)(short)(off & 0xffff), d), wrap(f2 - (off >> 16), d));
            }
          __syncwarp();
        }
        if (worker) {
#pragma unroll 4
          for (int k = rsub; k < K; k += RPP) {
            const int off = s_off[k];
            int4 *hp = hq + (wrap(f2 - (off >> 16), d) * d + wrap(f1 - (int)(short)(off & 0xffff), d)) * HP4 + part;
            int4 h = *hp;
            const int4 w = s_W[k * HP4 + part];
            h.x += sgn * w.x; h.y += sgn * w.y; h.z += sgn * w.z; h.w += sgn * w.w;
            *hp = h;
          }
        }
        if (GATED) {
          for (unsigned todo = mychg; todo; todo &= todo - 1) {
            const int off = s_off[32 * (__ffs(todo) - 1) + lane];
            const int row = wrap(f2 - (off >> 16), d) * d + wrap(f1 - (int)(short)(off & 0xffff), d);
            atomicXor(s_act + (row >> 5), 1u << (row & 31));
          }
          if (full) {       // persistent list: append the rows that switched on, swap-remove those that switched off
            for (unsigned lanes = __ballot_sync(SO_FULL, mychg != 0); lanes; lanes &= lanes - 1) {
              const int l = __ffs(lanes) - 1;
              const unsigned o = __shfl_sync(SO_FULL, myon, l);
              for (unsigned m = __shfl_sync(SO_FULL, mychg, l); m; m &= m - 1) {
                const int j = __ffs(m) - 1;
                const int off = s_off[32 * j + l];
                const int r1 = wrap(f1 - (int)(short)(off & 0xffff), d), r2 = wrap(f2 - (off >> 16), d);
                const int row = r2 * d + r1;
                const bool on = (o >> j) & 1u;
                if (lane == 0) {
                  if (on) {
                    s_list[n_alist] = (unsigned)row | ((unsigned)r1 << 16) | ((unsigned)r2 << 24);
                    s_pos[row] = (unsigned short)n_alist;
                  } else {
                    const int ppos = s_pos[row];
                    const unsigned last = s_list[n_alist - 1];
                    s_list[ppos] = last;
                    s_pos[last & 0xffffu] = (unsigned short)ppos;
                  }
                }
                n_alist += on ? 1 : -1;
                __syncwarp();
              }
            }
          }
        }
        if (lane == 0) s_bits[f >> 5] ^= (1u << (f & 31));
        if (GATED && __any_sync(SO_FULL, mywalk != 0)) {     // ... and enter the paths of the new structure
          __syncwarp();
          for (int j = 0; j < KW; ++j)
            if ((mywalk >> j) & 1) {
              const int off = s_off[32 * j + lane];
              walk_mark<true>(s_tree, s_bits, s_S, KW, d, wrap(f1 - (int)(short)(off & 0xffff), d), wrap(f2 - (off >> 16), d));
            }
        }
        lo += dlo_w;
        hi += dhi + (lo >> SO_SPLIT_BITS);
        lo &= SO_SPLIT_MASK;
        nact += dn;
        ++n_acc;
      }
      if (P.accept_out && lane == 0) P.accept_out[c * P.n_steps + s] = accept ? 1 : 0;
      __syncwarp();
    }
    for (int w = lane; w < W; w += 32) bits_g[w] = s_bits[w];
    if constexpr (RES)
      for (int q = lane; q < N * HP4; q += 32) hq_g[q] = s_hq[q];
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


// ---- few chains: one CTA (TEAM_WARPS warps) per chain -------------------------------------------------------------------
// The reference anneals ONE chain per process; with a warp per chain a launch of a few chains leaves the GPU idle and
// a step costs the latency of one warp walking all K rows.  Here a chain's step is spread over the warps of a CTA:
// the gate pass by row groups of 32 (one group per warp at K = 144), the difference pass by the same groups (gated) or
// by contiguous column chunks (open); every warp reduces its part exactly (REDUX limbs), the parts meet in shared
// memory at one __syncthreads per step, and every thread takes the same decision from the same sums.  Same caches
// (gate bits, path index) and same arithmetic as sa_resident_kernel: results are bit-identical.
constexpr int TEAM_WARPS = 4;

template <int HP4, bool GATED>
__global__ void __launch_bounds__(TEAM_WARPS * 32) sa_team_kernel(SaParams P) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  constexpr int RPP = 32 / HP4, STRIDE = RPP * HP4;
  const SurView &sv = P.sv;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = P.d, N = P.N, K = sv.K, KW = (K + 31) >> 5, W = P.W;
  // trees of up to SO_TREE_SMEM_NODES nodes are copied to shared memory, larger ones are walked in global memory (L1/L2)
  const int n_tree = (GATED && sv.n_tree <= SO_TREE_SMEM_NODES) ? sv.n_tree : 0;
  // shared layout: [W1q][tree][offsets][w2q][hq: N*HP4 int4][parts: 2 x TEAM_WARPS x 3 int64][chain id][bits W][act W]
  // [S: N*KW][lists: KW*32 u32][positions: K u16][counts: KW int]
  int4 *s_W = (int4 *)s_raw;
  int4 *s_tree_buf = s_W + (size_t)K * HP4;
  const int4 *s_tree = n_tree ? s_tree_buf : sv.t_packed;
  int32_t *s_off = (int32_t *)(s_tree_buf + n_tree);
  int4 *s_w2 = (int4 *)((unsigned char *)s_off + a16((size_t)K * 4));
  int4 *hq = s_w2 + HP4;
  long long *s_part = (long long *)(hq + (size_t)N * HP4);
  long long *s_chain = s_part + 2 * TEAM_WARPS * 3;
  uint32_t *s_bits = (uint32_t *)(s_chain + 1);
  uint32_t *s_act = s_bits + W;
  uint32_t *s_S = s_act + W;
  uint32_t *s_list = s_S + (GATED ? N * KW : 0);
  int *s_cnt = (int *)(s_list + (GATED ? KW * 32 : 0));
  unsigned short *s_pos = (unsigned short *)(s_cnt + KW);
  for (int q = threadIdx.x; q < K * HP4; q += blockDim.x) s_W[q] = ((const int4 *)sv.W1q)[q];
  for (int k = threadIdx.x; k < K; k += blockDim.x) s_off[k] = sv.offs[k];
  if (threadIdx.x < HP4) s_w2[threadIdx.x] = ((const int4 *)sv.w2q)[threadIdx.x];
  if (GATED)
    for (int i = threadIdx.x; i < n_tree; i += blockDim.x) s_tree_buf[i] = sv.t_packed[i];
  const int part = lane % HP4, rsub = lane / HP4;
  const bool worker = lane < STRIDE;
  const int4 w2 = worker ? __ldg((const int4 *)sv.w2q + part) : make_int4(0, 0, 0, 0);
  const float tol_f = (float)P.guard_tol;
  const bool guard_on = P.exact_guard != 0;
  const int chunk = (K + TEAM_WARPS - 1) / TEAM_WARPS;      // open: columns [warp * chunk, (warp + 1) * chunk)

  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) *s_chain = (long long)atomicAdd(P.counters, 1ULL);
    __syncthreads();
    const long long c = *s_chain;
    if (c >= P.n_chains) break;
    uint32_t *bits_g = P.bits + c * W;
    int4 *hq_g = (int4 *)P.hq + c * (long long)N * HP4;
    for (int w = threadIdx.x; w < W; w += blockDim.x) s_bits[w] = bits_g[w];
    for (int q = threadIdx.x; q < N * HP4; q += blockDim.x) hq[q] = hq_g[q];
    if (GATED)
      for (int i = threadIdx.x; i < N * KW; i += blockDim.x) s_S[i] = 0u;
    long long hi = P.hi[c], lo = P.lo[c];      // tracked identically by every thread
    int nact = P.nact[c];
    const double tsc = P.tscale[c];
    const float tsc_f = (float)tsc;
    unsigned n_acc = 0, n_guard = 0;
    __syncthreads();
    if (GATED) {
      for (int r0 = 32 * warp; r0 < N; r0 += 32 * TEAM_WARPS) {
        const int r = r0 + lane;
        int g = 0;
        if (r < N) {
          const int r2 = r / d;
          g = walk_mark<true>(s_tree, s_bits, s_S, KW, d, r - r2 * d, r2);
        }
        __syncwarp();
        const unsigned m = __ballot_sync(SO_FULL, g);
        if (lane == 0) s_act[r0 >> 5] = m;
      }
      __syncthreads();
    }
    DrawR batch;
    batch.f = 0; batch.u = 0.f; batch.tf = 0.f;

    for (long long s = 0; s < P.n_steps; ++s) {
      if ((s & 31) == 0)
        batch = draw_batch_res(P.temps, P.proposals, P.uniforms, P.n_steps, P.step0, P.chain_id0 + c, c, P.seed, N, d,
                               s >> 5, lane);
      const int fp = __shfl_sync(SO_FULL, batch.f, (int)(s & 31));
      const float u = __shfl_sync(SO_FULL, batch.u, (int)(s & 31));
      const float Tf = tsc_f * __shfl_sync(SO_FULL, batch.tf, (int)(s & 31));
      const int f1 = fp & 0xffff, f2 = fp >> 16;
      const int f = f2 * d + f1;
      const int sgn = ((s_bits[f >> 5] >> (f & 31)) & 1) ? -1 : 1;
      long long dhi = 0;
      int dlo = 0, dn = 0;
      unsigned mywalk = 0, mychg = 0;      // bit j: this lane's row of group j (groups j = warp, warp + TEAM_WARPS, ...)
      int ax = 0, ay = 0, az = 0, aw = 0;
      if (GATED) {
        unsigned g0m = 0;
        const unsigned lt = (1u << lane) - 1u;
        for (int j = warp; j < KW; j += TEAM_WARPS) {
          const int k = 32 * j + lane;
          const bool valid = k < K;
          const int off = s_off[valid ? k : 0];
          const int row = wrap(f2 - (off >> 16), d) * d + wrap(f1 - (int)(short)(off & 0xffff), d);
          const unsigned g0 = valid ? (s_act[row >> 5] >> (row & 31)) & 1u : 0u;
          g0m |= g0 << j;
          mywalk |= ((s_S[f * KW + j] >> lane) & 1u) << j;
          const unsigned m = __ballot_sync(SO_FULL, g0);
          if (g0) {
            const int pos = __popc(m & lt);
            s_list[32 * j + pos] = (unsigned)(k << 16) | 0x8000u | (unsigned)row;
            s_pos[k] = (unsigned short)pos;
          }
          if (lane == 0) s_cnt[j] = __popc(m);
        }
        unsigned g1m = g0m;
        for (unsigned todo = mywalk; todo; todo &= todo - 1) {
          const int j = __ffs(todo) - 1;
          const int off = s_off[32 * j + lane];
          const unsigned g1 = walk1(s_tree, s_bits, d, f, wrap(f1 - (int)(short)(off & 0xffff), d), wrap(f2 - (off >> 16), d));
          g1m = (g1m & ~(1u << j)) | (g1 << j);
        }
        __syncwarp();
        mychg = g0m ^ g1m;
        for (unsigned todo = mychg; todo; todo &= todo - 1) {
          const int j = __ffs(todo) - 1, k = 32 * j + lane;
          const int off = s_off[k];
          const int row = wrap(f2 - (off >> 16), d) * d + wrap(f1 - (int)(short)(off & 0xffff), d);
          const int on = (g1m >> j) & 1;
          if (!on) s_list[32 * j + s_pos[k]] &= ~0x8000u;
          const long long A = row_sum_res<HP4>(hq + row * HP4, s_W + k * HP4, s_w2, on ? sgn : 0);
          const int sg = on ? 1 : -1;
          dhi += sg * (A >> SO_SPLIT_BITS);
          dlo += sg * (int)(A & SO_SPLIT_MASK);
          dn += sg;
        }
        __syncwarp();
        if (worker) {
          for (int j = warp; j < KW; j += TEAM_WARPS) {
            const int cnt = s_cnt[j];
            for (int i = rsub; i < cnt; i += RPP) {
              const unsigned e = s_list[32 * j + i];
              const int sg = sgn & -(int)((e >> 15) & 1u);
              const int4 h = hq[(e & 0x7fffu) * HP4 + part], w = s_W[(e >> 16) * HP4 + part];
              ax += max(h.x + sg * w.x, 0) - max(h.x, 0);
              ay += max(h.y + sg * w.y, 0) - max(h.y, 0);
              az += max(h.z + sg * w.z, 0) - max(h.z, 0);
              aw += max(h.w + sg * w.w, 0) - max(h.w, 0);
            }
          }
        }
      } else if (worker) {
        const int k_hi = min(K, (warp + 1) * chunk);
#pragma unroll 4
        for (int k = warp * chunk + rsub; k < k_hi; k += RPP) {
          const int off = s_off[k];
          const int row = wrap(f2 - (off >> 16), d) * d + wrap(f1 - (int)(short)(off & 0xffff), d);
          const int4 h = hq[row * HP4 + part], w = s_W[k * HP4 + part];
          ax += max(h.x + sgn * w.x, 0) - max(h.x, 0);
          ay += max(h.y + sgn * w.y, 0) - max(h.y, 0);
          az += max(h.z + sgn * w.z, 0) - max(h.z, 0);
          aw += max(h.w + sgn * w.w, 0) - max(h.w, 0);
        }
      }
      long long acc = (long long)w2.x * ax + (long long)w2.y * ay + (long long)w2.z * az + (long long)w2.w * aw;
      {
        int l0 = (int)(acc & 0x1fffff), l1 = (int)((acc >> 21) & 0x1fffff), l2 = (int)(acc >> 42);
        l0 = __reduce_add_sync(SO_FULL, l0);
        l1 = __reduce_add_sync(SO_FULL, l1);
        l2 = __reduce_add_sync(SO_FULL, l2);
        acc = ((long long)l2 << 42) + ((long long)l1 << 21) + (long long)l0;
      }
      long long dlo_w = acc & SO_SPLIT_MASK;
      if (GATED && __any_sync(SO_FULL, mychg != 0)) {
        int l0 = (int)(dhi & 0x1fffff), l1 = (int)((dhi >> 21) & 0x1fffff), l2 = (int)(dhi >> 42);
        l0 = __reduce_add_sync(SO_FULL, l0);
        l1 = __reduce_add_sync(SO_FULL, l1);
        l2 = __reduce_add_sync(SO_FULL, l2);
        dhi = ((long long)l2 << 42) + ((long long)l1 << 21) + (long long)l0;
        dn = __reduce_add_sync(SO_FULL, dn);
        dlo_w += (long long)__reduce_add_sync(SO_FULL, dlo);
      }
      dhi += acc >> SO_SPLIT_BITS;
      // the warps' parts meet here (double-buffered by step parity: one barrier per step is enough)
      long long *mine = s_part + ((s & 1) * TEAM_WARPS + warp) * 3;
      if (lane == 0) { mine[0] = dhi; mine[1] = dlo_w; mine[2] = dn; }
      __syncthreads();
      dhi = 0; dlo_w = 0; dn = 0;
#pragma unroll
      for (int w = 0; w < TEAM_WARPS; ++w) {
        const long long *o = s_part + ((s & 1) * TEAM_WARPS + w) * 3;
        dhi += o[0]; dlo_w += o[1]; dn += (int)o[2];
      }
      const double delta = so_real(sv.c, sv.b2, sv.y_scale, sv.y_mean, dhi, dlo_w, dn);
      bool accept;
      {
        const float pf = __expf(__fdividef((float)delta, Tf));
        if (delta > 0.0 && (!guard_on || delta >= P.guard_tol)) {     // (with the guard: only when no tie is possible)
          accept = true;
        } else if (delta <= 0.0 && Tf > 1e-30f && u > 0.f && fabsf(pf - u) > 1e-3f * fmaxf(pf, u) &&
                   (!guard_on || tol_f < 5e-4f * Tf)) {
          accept = pf > u;
        } else {
          int r = decide_res(delta, tsc, P.temps, s, u, P.exact_guard, P.guard_tol);
          if (r & 2) {
            r = exact_decision_res(sv, s_bits, d, f, lane, tsc, P.temps, s, u);     // every warp, same result
            ++n_guard;
          }
          accept = r & 1;
        }
      }
      if (accept) {                                  // uniform over the CTA
        if (P.type_hist && warp == 0) so_track_flip(P.type_hist, P.n_chains, c, s_bits, d, N, f, lane);   // before the bit flips
        if (GATED) {
          for (unsigned todo = mywalk; todo; todo &= todo - 1) {
            const int off = s_off[32 * (__ffs(todo) - 1) + lane];
            walk_mark<false>(s_tree, s_bits, s_S, KW, d, wrap(f1 - (int)(short)(off & 0xffff), d), wrap(f2 - (off >> 16), d));
          }
          for (unsigned todo = mychg; todo; todo &= todo - 1) {
            const int off = s_off[32 * (__ffs(todo) - 1) + lane];
            const int row = wrap(f2 - (off >> 16), d) * d + wrap(f1 - (int)(short)(off & 0xffff), d);
            atomicXor(s_act + (row >> 5), 1u << (row & 31));
          }
        }
        if (worker) {
          const int k_hi = min(K, (warp + 1) * chunk);
          for (int k = warp * chunk + rsub; k < k_hi; k += RPP) {
            const int off = s_off[k];
            int4 *hp = hq + (wrap(f2 - (off >> 16), d) * d + wrap(f1 - (int)(short)(off & 0xffff), d)) * HP4 + part;
            int4 h = *hp;
            const int4 w = s_W[k * HP4 + part];
            h.x += sgn * w.x; h.y += sgn * w.y; h.z += sgn * w.z; h.w += sgn * w.w;
            *hp = h;
          }
        }
        __syncthreads();                             // old paths dropped everywhere before the bit flips
        if (threadIdx.x == 0) s_bits[f >> 5] ^= (1u << (f & 31));
        __syncthreads();
        if (GATED) {
          for (unsigned todo = mywalk; todo; todo &= todo - 1) {
            const int off = s_off[32 * (__ffs(todo) - 1) + lane];
            walk_mark<true>(s_tree, s_bits, s_S, KW, d, wrap(f1 - (int)(short)(off & 0xffff), d), wrap(f2 - (off >> 16), d));
          }
          __syncthreads();
        }
        lo += dlo_w;
        hi += dhi + (lo >> SO_SPLIT_BITS);
        lo &= SO_SPLIT_MASK;
        nact += dn;
        ++n_acc;
      }
      if (P.accept_out && threadIdx.x == 0) P.accept_out[c * P.n_steps + s] = accept ? 1 : 0;
    }
    __syncthreads();
    for (int w = threadIdx.x; w < W; w += blockDim.x) bits_g[w] = s_bits[w];
    for (int q = threadIdx.x; q < N * HP4; q += blockDim.x) hq_g[q] = hq[q];
    if (threadIdx.x == 0) {
      P.hi[c] = hi;
      P.lo[c] = lo;
      P.nact[c] = nact;
      if (n_acc) atomicAdd(P.counters + 1, (unsigned long long)n_acc);
      if (n_guard) atomicAdd(P.counters + 2, (unsigned long long)n_guard);
    }
  }
}

}  // namespace

static size_t shared_tables_bytes(const so_surrogate *s) {
  const int n_tree = (s->gated && s->n_tree <= SO_TREE_SMEM_NODES) ? s->n_tree : 0;
  return (size_t)s->K * s->HP4 * 16 + (size_t)n_tree * 16 + a16((size_t)s->K * 4) + (size_t)s->HP4 * 16;
}
static bool shape_ok(const so_surrogate *s, int N) {
  return !(s->K > 1024 || s->HP4 > 8 || N > 32767);
}

// Warps per CTA and its shared memory, maximising the resident chains of an SM; false if the state does not fit
// (or the tree is too large to share).
bool so_resident_plan(const so_surrogate *s, int N, int W, int *warps_out, size_t *smem_out, int min_warps_per_sm,
                      int *index_global_out) {
  const size_t kSmemSM = 228 * 1024, kSmemCTA = 227 * 1024;
  const int KW = (s->K + 31) / 32;
  if (!shape_ok(s, N)) return false;
  const size_t shared = shared_tables_bytes(s);
  int best_w = 0, best_total = 0, best_glob = 0;
  size_t best_smem = 0;
  const int cand[] = {4, 5, 6, 7, 8, 10, 12, 14, 16};
  // gated: the path index (N * KW words per warp) in shared memory, or in the per-warp global scratch when that fits more
  // warps on an SM (at most 16: the kernel needs its 128 registers per thread)
  const char *env = getenv("SO_RES_INDEX_GLOBAL");
  for (int glob = 0; glob <= (s->gated ? 1 : 0); ++glob) {
    if (env && s->gated && (env[0] == '1') != (glob == 1)) continue;
    const size_t per_warp = (size_t)N * s->HP4 * 16 +
                            a16((size_t)(2 * W + (s->gated ? (glob ? 0 : N * KW) + s->K : 0)) * 4 + (s->gated ? (size_t)s->K * 2 : 0));
    for (int w : cand) {
      size_t smem = shared + (size_t)w * per_warp;
      if (smem > kSmemCTA) continue;
      int per_sm = (int)(kSmemSM / (smem + 1024));
      if (per_sm * w > 64) per_sm = 64 / w;
      int total = per_sm * w;
      if (s->gated && total > 16) total = 16;       // (register file: 16 warps of 128 registers)
      if (total > best_total) { best_total = total; best_w = w; best_smem = smem; best_glob = glob; }
    }
  }
  if (best_total < min_warps_per_sm) return false;   // too few resident chains to keep an SM busy: the global-memory kernels win
  if (warps_out) *warps_out = best_w;
  if (smem_out) *smem_out = best_smem;
  if (index_global_out) *index_global_out = best_glob;
  return true;
}

// few chains on a small lattice: one CTA per chain (sa_team_kernel)
static size_t team_smem_bytes(const so_surrogate *s, int N, int W) {
  const int KW = (s->K + 31) / 32;
  size_t b = shared_tables_bytes(s) + (size_t)N * s->HP4 * 16 + (size_t)(2 * W) * 4;
  if (s->gated) b += (size_t)(N * KW + KW * 32) * 4 + a16((size_t)s->K * 2);
  return b + a16((size_t)KW * 4) + 2 * TEAM_WARPS * 3 * 8 + 8 + 64;
}
bool so_team_ok(const so_surrogate *s, int N, int W, long long n_chains, int n_sm) {
  // pays only while the SMs are otherwise idle: measured break-even at about two chains per SM (the four warps of a
  // team repeat the draws and the decision, so a team costs roughly twice the issue slots of a single warp)
  // (up to one CTA per SM with all of its shared memory: a single chain on a 48x48 lattice still runs from shared memory)
  const size_t smem = team_smem_bytes(s, N, W);
  if (!shape_ok(s, N) || smem > 227 * 1024) return false;
  return n_chains <= (smem <= 113 * 1024 ? 2LL : 1LL) * n_sm;
}

int so_launch_anneal_team(const so_chains *c, const SaParams &P, cudaStream_t st) {
  const so_surrogate *s = c->sur;
  const size_t smem = team_smem_bytes(s, P.N, P.W);
  const int blocks = (int)P.n_chains, warps = TEAM_WARPS;
#define SO_TEAM_LAUNCH(V, G)                                                                               \
  {                                                                                                        \
    auto kern = sa_team_kernel<V, G>;                                                                      \
    SO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
    kern<<<blocks, warps * 32, smem, st>>>(P);                                                             \
  }
#define SO_RES_CASE(V) \
  case V:              \
    if (s->gated) SO_TEAM_LAUNCH(V, true) else SO_TEAM_LAUNCH(V, false) break;
  switch (s->HP4) {
    SO_RES_CASE(1) SO_RES_CASE(2) SO_RES_CASE(3) SO_RES_CASE(4) SO_RES_CASE(5) SO_RES_CASE(6) SO_RES_CASE(7) SO_RES_CASE(8)
    default: so_set_error("hidden width not supported"); return SO_EINVAL;
  }
#undef SO_RES_CASE
#undef SO_TEAM_LAUNCH
  SO_CUDA(cudaGetLastError());
  return SO_OK;
}

// gated surrogate on a lattice too large for shared memory: can the cached-gate kernel (global state) run it?
bool so_gate_index_ok(const so_surrogate *s, int N, int W) {
  if (!s->gated || !shape_ok(s, N)) return false;
  const size_t per_warp = a16((size_t)(2 * W + s->K) * 4 + (size_t)s->K * 2);
  return shared_tables_bytes(s) + 8 * per_warp <= 100 * 1024;
}

#define SO_RES_LAUNCH(V, G, R)                                                                             \
  {                                                                                                        \
    auto kern = sa_resident_kernel<V, G, R>;                                                               \
    SO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
    kern<<<blocks, warps * 32, smem, st>>>(P);                                                             \
  }

int so_launch_anneal_resident(const so_chains *c, const SaParams &P0, cudaStream_t st) {
  const so_surrogate *s = c->sur;
  int n_sm = 0, warps = 0, index_global = 0;
  size_t smem = 0;
  SO_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, c->device));
  if (!so_resident_plan(s, P0.N, P0.W, &warps, &smem, 1, &index_global)) { so_set_error("resident kernel: state does not fit"); return SO_EINVAL; }
  const int per_sm = (int)((228 * 1024) / (smem + 1024));
  long long want = (P0.n_chains + warps - 1) / warps;
  int blocks = (int)(want < (long long)n_sm * per_sm ? want : (long long)n_sm * per_sm);
  if (blocks < 1) blocks = 1;
  SaParams P = P0;
  if (index_global) {                              // path index of every resident warp in a scratch buffer of the handle
    const size_t need = (size_t)blocks * warps * P.N * ((s->K + 31) / 32) * 4;
    if (c->gate_scratch_bytes < need) {
      if (c->gate_scratch) cudaFree(c->gate_scratch);
      c->gate_scratch = nullptr;
      c->gate_scratch_bytes = 0;
      if (cudaMalloc(&c->gate_scratch, need) != cudaSuccess) {
        cudaGetLastError();
        so_set_error("out of device memory for the gate index (%zu bytes)", need);
        return SO_ENOMEM;
      }
      c->gate_scratch_bytes = need;
    }
    P.gate_scratch = c->gate_scratch;
    P.variant_flags |= 8;
  }
#define SO_RES_CASE(V) \
  case V:              \
    if (s->gated) SO_RES_LAUNCH(V, true, true) else SO_RES_LAUNCH(V, false, true) break;
  switch (s->HP4) {
    SO_RES_CASE(1) SO_RES_CASE(2) SO_RES_CASE(3) SO_RES_CASE(4) SO_RES_CASE(5) SO_RES_CASE(6) SO_RES_CASE(7) SO_RES_CASE(8)
    default: so_set_error("hidden width not supported"); return SO_EINVAL;
  }
#undef SO_RES_CASE
  SO_CUDA(cudaGetLastError());
  return SO_OK;
}

// gated, large lattice: 8 warps per CTA, four CTAs per SM (64 registers per thread: the step is a chain of dependent
// global loads, occupancy pays more than the spills cost); the path index of each resident
// warp lives in a scratch buffer of the chains handle
int so_launch_anneal_gate_index(const so_chains *c, const SaParams &P0, cudaStream_t st) {
  const so_surrogate *s = c->sur;
  int n_sm = 0;
  const int warps = 8;
  SO_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, c->device));
  if (!so_gate_index_ok(s, P0.N, P0.W)) { so_set_error("cached-gate kernel: shape not supported"); return SO_EINVAL; }
  const int KW = (s->K + 31) / 32;
  const size_t smem = shared_tables_bytes(s) + warps * a16((size_t)(2 * P0.W + s->K) * 4 + (size_t)s->K * 2);
  long long want = (P0.n_chains + warps - 1) / warps;
  int blocks = (int)(want < (long long)n_sm * 4 ? want : (long long)n_sm * 4);
  if (blocks < 1) blocks = 1;
  const size_t need = (size_t)blocks * warps * P0.N * KW * 4;
  if (c->gate_scratch_bytes < need) {
    if (c->gate_scratch) cudaFree(c->gate_scratch);
    c->gate_scratch = nullptr;
    c->gate_scratch_bytes = 0;
    if (cudaMalloc(&c->gate_scratch, need) != cudaSuccess) {
      cudaGetLastError();
      so_set_error("out of device memory for the gate index (%zu bytes)", need);
      return SO_ENOMEM;
    }
    c->gate_scratch_bytes = need;
  }
  SaParams P = P0;
  P.gate_scratch = c->gate_scratch;
#define SO_RES_CASE(V) \
  case V:              \
    SO_RES_LAUNCH(V, true, false) break;
  switch (s->HP4) {
    SO_RES_CASE(1) SO_RES_CASE(2) SO_RES_CASE(3) SO_RES_CASE(4) SO_RES_CASE(5) SO_RES_CASE(6) SO_RES_CASE(7) SO_RES_CASE(8)
    default: so_set_error("hidden width not supported"); return SO_EINVAL;
  }
#undef SO_RES_CASE
  SO_CUDA(cudaGetLastError());
  return SO_OK;
}
