// so_types.cuh -- closed-form site-typing rules on the bit-packed layer-3 occupancy, shared by the
// descriptor kernels (so_desc.cu) and the SA kernels that carry type histograms through accepted flips (track_types).
//   LSC site typing   OML/LSC_cat.py:176-220        NH3 site typing (19 types)   OML/NH3_cat.py:169-396, export map :421
#pragma once
#include "so_internal.cuh"

// ring order: consecutive entries are mutually adjacent
static __device__ __constant__ int8_t c_ring1[6] = {1, 0, -1, -1, 0, 1};
static __device__ __constant__ int8_t c_ring2[6] = {0, 1, 1, 0, -1, -1};

// ---- closed-form typing rules -----------------------------------------------------------------------
struct L3Info {
  int x, c6, adj;   // occupancy, Ni neighbours, "two ring-adjacent vacancies"
};

__device__ __forceinline__ L3Info l3_info(const BitView &b, int i1, int i2) {
  L3Info r;
  r.x = b.at(i1, i2);
  int ring = 0;
#pragma unroll
  for (int k = 0; k < 6; ++k) ring |= b.at(i1 + c_ring1[k], i2 + c_ring2[k]) << k;
  r.c6 = __popc(ring);
  int vac = (~ring) & 63;
  int rot = ((vac << 1) | (vac >> 5)) & 63;
  r.adj = (vac & rot) != 0;
  return r;
}
__device__ __forceinline__ int lsc_type_of(const L3Info &s) {
  if (!s.x) return 0;
  if (s.c6 == 6) return 1;
  if (s.c6 == 4 && s.adj) return 2;
  return 0;
}
// NH3 base type of a layer-3 site: Ni -> 3 / 5 / 4, vacancy -> 6 (refined by l3_full)
__device__ __forceinline__ int nh3_t3_of(const L3Info &s) {
  if (!s.x) return 6;
  if (s.c6 == 6) return 3;
  if (s.c6 == 4 && s.adj) return 5;
  return 4;
}
// number of Ni under layer-4 atom (j1,j2): sites (0,-1),(1,-1),(0,0)
__device__ __forceinline__ int n_tri4(const BitView &b, int j1, int j2) {
  return b.at(j1, j2 - 1) + b.at(j1 + 1, j2 - 1) + b.at(j1, j2);
}
// number of Ni over layer-2 atom (k1,k2): sites (1,-1),(0,0),(1,0)
__device__ __forceinline__ int n_tri2(const BitView &b, int k1, int k2) {
  return b.at(k1 + 1, k2 - 1) + b.at(k1, k2) + b.at(k1 + 1, k2);
}
__device__ __forceinline__ int nh3_l4_type(const BitView &b, int j1, int j2) {
  int n = n_tri4(b, j1, j2);
  if (n == 1) return 12;
  if (n == 2) return 14;
  if (n == 0) return 0;
  int e = 0;
  int t;
  t = nh3_t3_of(l3_info(b, j1, j2 - 1)); e += (t == 4 || t == 5);
  t = nh3_t3_of(l3_info(b, j1 + 1, j2 - 1)); e += (t == 4 || t == 5);
  t = nh3_t3_of(l3_info(b, j1, j2)); e += (t == 4 || t == 5);
  return e == 0 ? 1 : (e == 1 ? 17 : 16);
}
__device__ __forceinline__ int nh3_l2_type(const BitView &b, int k1, int k2) {
  int n = n_tri2(b, k1, k2);
  if (n == 0) return 8;
  if (n == 2) return 15;
  if (n == 1) return 0;
  int ntop = 0, nedge = 0, t;
  t = nh3_t3_of(l3_info(b, k1 + 1, k2 - 1)); ntop += (t == 3); nedge += (t == 5);
  t = nh3_t3_of(l3_info(b, k1, k2)); ntop += (t == 3); nedge += (t == 5);
  t = nh3_t3_of(l3_info(b, k1 + 1, k2)); ntop += (t == 3); nedge += (t == 5);
  if (ntop == 2 && nedge == 1) return 18;
  if (ntop == 1 && nedge == 2) return 19;
  return 2;
}
// full layer-3 type incl. the vacancy refinements 10 / 11 / 9
__device__ __forceinline__ int nh3_l3_type(const BitView &b, int i1, int i2, const L3Info &s, const int32_t *h5_ptr,
                                           const int32_t *h5_idx) {
  int t = nh3_t3_of(s);
  if (t != 6) return t;
  // layer-4 atoms above: (0,0), (-1,1), (0,1)
  int na = n_tri4(b, i1, i2), nb = n_tri4(b, i1 - 1, i2 + 1), nc = n_tri4(b, i1, i2 + 1);
  if (na == 2 || nb == 2 || nc == 2) return 11;
  if (na == 1 || nb == 1 || nc == 1) return 10;
  int d = b.d;
  int v = i2 * d + i1;
  for (int p = h5_ptr[v]; p < h5_ptr[v + 1]; ++p) {
    int j = h5_idx[p];
    int k2 = j / d, k1 = j - k2 * d;
    if (n_tri2(b, k1, k2) == 2) return 9;
  }
  return 6;
}
__device__ __forceinline__ int nh3_l1_type(const BitView &b, int m1, int m2) {
  // layer-2 atoms above: (-1,0), (0,-1), (0,0); type 8 <=> n_tri2 == 0, None <=> n_tri2 == 1
  int a = n_tri2(b, m1 - 1, m2), c = n_tri2(b, m1, m2 - 1), e = n_tri2(b, m1, m2);
  int n8 = (a == 0) + (c == 0) + (e == 0);
  int n0 = (a == 1) + (c == 1) + (e == 1);
  if (n8 == 3) return 7;
  if (n8 == 2 && n0 == 1) return 13;
  return 0;
}
__device__ __forceinline__ int nh3_export_code(int t) {
  // {1:1, 3:2, 5:3, 16:4, 17:5, 18:6, 19:7}  (OML/NH3_cat.py:421)
  switch (t) {
    case 1: return 1;
    case 3: return 2;
    case 5: return 3;
    case 16: return 4;
    case 17: return 5;
    case 18: return 6;
    case 19: return 7;
    default: return 0;
  }
}

// ---- per-flip histogram deltas: one affected node per lane ---------------------------------------------------------
// Nodes whose exported type can change when layer-3 site f flips, as (layer, offset from f): the 7 layer-3 sites within
// one ring, then the layer-4 / layer-2 atoms whose triangle touches one of them (ring x {(0,0),(-1,1),(0,1)} and
// ring x {(-1,0),(0,0),(-1,1)}, duplicates removed): 7 + 12 + 12 = 31 entries.
#define SO_AFF_N 31
static __device__ __constant__ int8_t c_aff_layer[SO_AFF_N] = {3, 3, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4,
                                                              4, 4, 4, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2};
static __device__ __constant__ int8_t c_aff_o1[SO_AFF_N] = {0, 1, 0, -1, -1, 0, 1, 0, -1, 0, 1, 1, -1, 0, -2, -1,
                                                           -2, 0, 1, -1, 0, -1, 1, 0, -1, -2, -2, -2, -1, 0, 1};
static __device__ __constant__ int8_t c_aff_o2[SO_AFF_N] = {0, 0, 1, 1, 0, -1, -1, 0, 1, 1, 0, 1, 2, 2, 2, 0,
                                                           1, -1, -1, 0, 0, 1, 0, 1, 2, 1, 2, 0, -1, -1, -1};

// Histogram deltas of flipping site f of the structure `w` (as it is BEFORE the flip), whole warp: returns in lane t
// (t < 8) the change of exported-NH3-type count t and in lane 8 + t (t < 3) the change of LSC-type count t; other
// lanes 0.  The rule of flip_delta_kernel, callable from the SA kernels' accept path.
static __device__ __noinline__ int so_flip_hist_delta(const uint32_t *w, int d, int N, int f, int lane) {
  const int f2 = f / d, f1 = f - f2 * d;
  int code0 = -1, code1 = -1, lsc0 = -1, lsc1 = -1;
  int node = -1 - lane;   // unique negative id for idle lanes
  if (lane < SO_AFF_N) {
    int layer = c_aff_layer[lane];
    int i1 = wrap(f1 + c_aff_o1[lane], d), i2 = wrap(f2 + c_aff_o2[lane], d);
    node = layer * N + i2 * d + i1;
  }
  // aliasing on tiny lattices: only the lowest lane of equal nodes counts
  unsigned same = __match_any_sync(SO_FULL, node);
  bool leader = (lane < SO_AFF_N) && ((__ffs(same) - 1) == lane);
  if (leader) {
    int layer = c_aff_layer[lane];
    int v = node - layer * N;
    int i2 = v / d, i1 = v - i2 * d;
    BitView b0{w, d, -1}, b1{w, d, f};
    if (layer == 3) {
      L3Info s0 = l3_info(b0, i1, i2), s1 = l3_info(b1, i1, i2);
      lsc0 = lsc_type_of(s0); lsc1 = lsc_type_of(s1);
      code0 = nh3_export_code(nh3_t3_of(s0)); code1 = nh3_export_code(nh3_t3_of(s1));
    } else if (layer == 4) {
      code0 = nh3_export_code(nh3_l4_type(b0, i1, i2)); code1 = nh3_export_code(nh3_l4_type(b1, i1, i2));
    } else {
      code0 = nh3_export_code(nh3_l2_type(b0, i1, i2)); code1 = nh3_export_code(nh3_l2_type(b1, i1, i2));
    }
  }
  int out = 0;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    int dlt = __popc(__ballot_sync(SO_FULL, code1 == t)) - __popc(__ballot_sync(SO_FULL, code0 == t));
    if (lane == t) out = dlt;
  }
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    int dlt = __popc(__ballot_sync(SO_FULL, lsc1 == t)) - __popc(__ballot_sync(SO_FULL, lsc0 == t));
    if (lane == 8 + t) out = dlt;
  }
  return out;
}

// accept path of the SA kernels: the tracked histograms of chain `chain` += delta.  type_hist = [n_chains][8] exported
// NH3 counts followed by [n_chains][3] LSC counts.  Call BEFORE the bit of site f is flipped, all 32 lanes converged.
static __device__ __forceinline__ void so_track_flip(int32_t *type_hist, long long n_chains, long long chain,
                                                     const uint32_t *w, int d, int N, int f, int lane) {
  const int dlt = so_flip_hist_delta(w, d, N, f, lane);
  if (dlt) {
    if (lane < 8) type_hist[chain * 8 + lane] += dlt;
    else if (lane < 11) type_hist[n_chains * 8 + chain * 3 + (lane - 8)] += dlt;
  }
}
