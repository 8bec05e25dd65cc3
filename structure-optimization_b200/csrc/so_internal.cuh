// so_internal.cuh -- shared declarations of the sm_100a hot path (not part of the public C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/so_b200.h"

#define SO_SPLIT_BITS 20
#define SO_SPLIT_MASK ((1LL << SO_SPLIT_BITS) - 1)
#define SO_FULL 0xffffffffu

void so_set_error(const char *fmt, ...);
#define SO_CUDA(call)                                                                           \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      so_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));      \
      return SO_ECUDA;                                                                          \
    }                                                                                           \
  } while (0)

#define SO_REQUIRE(cond, ...)        \
  do {                               \
    if (!(cond)) {                   \
      so_set_error(__VA_ARGS__);     \
      return SO_EINVAL;              \
    }                                \
  } while (0)
#define SO_TRY(call)                 \
  do {                               \
    int rc__ = (call);               \
    if (rc__ != SO_OK) return rc__;  \
  } while (0)

// RAII device buffer for call-scoped staging
struct DevBuf {
  void *p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t bytes) {
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
      p = nullptr;
      so_set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
      return SO_ENOMEM;
    }
    return SO_OK;
  }
  template <class T> T *as() { return (T *)p; }
};


struct so_lattice {
  int d, N, W, kind, device;
  std::vector<int32_t> rowptr, col;     // host CSR of the 5N-node graph
  std::vector<double> pos;              // raw positions [5N][3]
  double cell[4];
  int32_t *h5_ptr_dev, *h5_idx_dev;     // NH3 type-9 pair table (layer-3 site -> layer-2 atoms, non-periodic)
};

struct so_surrogate {
  so_lattice *lat;
  int device;
  int K, H, Hp, HP4, e_h, e_w, gated, n_tree;
  double c, b2, y_mean, y_scale;
  int32_t *offs_dev;                    // [K] packed (o2 << 16) | (o1 & 0xffff)
  int32_t *W1q_dev, *b1q_dev, *w2q_dev; // [K][Hp], [Hp], [Hp]
  double *W1d_dev, *b1d_dev, *w2d_dev;  // fp64 originals [K][H], [H], [H] (exact guard)
  int32_t *t_feature_dev, *t_left_dev, *t_right_dev;
  double *t_thr_dev;
  uint8_t *t_active_dev;
  int4 *t_packed_dev;                   // one record per node: {offset of the tested column (o1 | o2<<16), left, right, flags}
  std::vector<int32_t> W1q_host, b1q_host, w2q_host, offs_host;
  int win_ok, win_a1, win_b1, win_a2, win_b2;   // descriptor offsets form the full rectangle [a1,b1] x [a2,b2]
  int32_t *Wst_dev;                             // W1q rows in staged (memory) order of the window
  int compact;                                  // 24-bit packed state, 64 bytes per site (so_surrogate_create_compact)
  int32_t *bias24_dev;                          // compact: [Hp] stored value = hq + bias
  int32_t *Wst24_dev;                           // compact: [K*4 slots][8] = five weights of the slot's hidden units + pad
  void *wtree_dev;                              // gated window kernel: 8-byte tree nodes (so_sa_window_gated.cu) or NULL
  unsigned char *tc_image_dev;                  // B operand of the tensor-core scorer (so_score_tc.cu) or NULL
  int tc_ok;
};

struct so_chains {
  so_lattice *lat;
  int device;
  so_surrogate *sur;
  int64_t n_chains;
  uint32_t *bits_dev;                   // [n_chains][W]
  int32_t *hq_dev;                      // [n_chains][N][Hp]
  long long *hi_dev, *lo_dev;           // objective sums: hi*2^20 + lo
  int32_t *nact_dev;                    // active sites (gated) or N
  double *tscale_dev;
  unsigned long long *counters_dev;     // [0]=work counter [1]=accepted [2]=guard fired [3]=invalid flag [4]=swaps
  double *xch_scratch;                  // exchange scratch [xch_cap]
  int64_t xch_cap;
  cudaStream_t stream;
  cudaEvent_t ev0, ev1;
  double *temps_dev;                    // schedule staging of so_anneal (grow-only: many short launches per optimize())
  size_t temps_cap;
  mutable uint32_t *gate_scratch;       // path index of the resident warps of the cached-gate kernel (so_sa_resident.cu)
  mutable size_t gate_scratch_bytes;
  int32_t *tmp_hq_dev;                  // compact state: int32 staging slab of the state build
  size_t tmp_hq_chains;
  int32_t *type_hist_dev;               // tracked histograms: [n_chains][8] NH3 exported + [n_chains][3] LSC, or null
  CUtensorMap tmap;                     // hq as int32 [n_chains][d][d*Hp], box = one window (so_sa_window.cu)
  int tmap_ok;
  void *gate_meta_dev;                  // gated window kernel: one 16-byte record per row {T lo, T hi | gate << 31, 0, 0}
  CUtensorMap tmap_meta;                // the records as int32 [n_chains][d][d*4], box = one window
  int tmap_meta_ok;
  mutable int meta_valid;               // the records describe the current bits (set by the gated window kernel, cleared by
                                        // everything else that changes bits)
};

// ---- device-side views ------------------------------------------------------------------------------
struct SurView {
  int K, Hp, HP4, gated;
  double c, b2, y_mean, y_scale;
  const int32_t *offs;
  const int32_t *W1q, *b1q, *w2q;
  const double *W1d, *b1d, *w2d;
  int H;
  const int32_t *t_feature, *t_left, *t_right;
  const double *t_thr;
  const uint8_t *t_active;
  const int32_t *bias24;                // compact state: per-unit storage bias
  const int4 *t_packed;                 // flags: 1 leaf, 2 leaf is active, 4 go left when the bit is 0, 8 ... when it is 1; bits 8.. = column tested
  int n_tree;
};

struct SaParams {
  int d, N, W;
  long long n_chains;
  uint32_t *bits;
  int32_t *hq;
  long long *hi, *lo;
  int32_t *nact;
  const double *tscale;
  const double *temps;                  // [n_steps] device
  long long step0, n_steps;
  unsigned long long seed;
  long long chain_id0;
  const int32_t *proposals;             // device or null
  const float *uniforms;
  uint8_t *accept_out;
  int exact_guard;
  double guard_tol;
  int variant;                          // 0 auto, 1 row-per-lane (gated) kernel, 2 generic flat kernel, 3 auto without the resident kernel,
                                        // 4 resident warp-per-chain kernel even for few chains, 5 CTA-per-chain kernel
  unsigned long long *counters;
  uint32_t *gate_scratch;
  int variant_flags;                    // bit 0: resident kernel takes a lane per whole row in the difference pass; bit 1: persistent
                                        // list of active rows; bit 3: path index of the resident kernel in the global scratch
  int32_t *type_hist;                   // [n_chains][11] tracked type histograms (so_track_types) or null
  SurView sv;
};

inline SurView make_view(const so_surrogate *s) {
  SurView v;
  v.K = s->K; v.Hp = s->Hp; v.HP4 = s->HP4; v.gated = s->gated; v.H = s->H;
  v.c = s->c; v.b2 = s->b2; v.y_mean = s->y_mean; v.y_scale = s->y_scale;
  v.bias24 = s->bias24_dev;
  v.offs = s->offs_dev; v.W1q = s->W1q_dev; v.b1q = s->b1q_dev; v.w2q = s->w2q_dev;
  v.W1d = s->W1d_dev; v.b1d = s->b1d_dev; v.w2d = s->w2d_dev;
  v.t_feature = s->t_feature_dev; v.t_left = s->t_left_dev; v.t_right = s->t_right_dev;
  v.t_thr = s->t_thr_dev; v.t_active = s->t_active_dev; v.t_packed = s->t_packed_dev;
  v.n_tree = s->n_tree;
  return v;
}

// ---- launch wrappers (defined in the kernel translation units) ----------------------------------------
int so_launch_pack(const uint8_t *occ_dev, uint32_t *bits_dev, int64_t n, int N, int W, unsigned long long *flag,
                   cudaStream_t st);
int so_launch_unpack(const uint32_t *bits_dev, uint8_t *occ_dev, int64_t n, int N, int W, cudaStream_t st);
int so_launch_randomize(uint32_t *bits_dev, int64_t n, int N, int W, uint64_t seed, double coverage, cudaStream_t st);
int so_launch_descriptors(const so_lattice *lat, const uint32_t *bits_dev, int64_t n, uint8_t *c6, uint8_t *lsc_type,
                          uint8_t *nh3_type, int32_t *lsc_hist, int32_t *nh3_hist, int32_t *nh3_exp, cudaStream_t st);
int so_launch_flip_delta(const so_lattice *lat, const uint32_t *bits_dev, int64_t n, const int32_t *sites_dev,
                         int32_t *d_lsc, int32_t *d_nh3_exp, cudaStream_t st);
int so_launch_score(const so_lattice *lat, const so_surrogate *s, const uint32_t *bits_dev, int64_t n, int32_t *hq_dev,
                    long long *hi, long long *lo, int32_t *nact, cudaStream_t st);
int so_normalize_sums(long long *hi, long long *lo, int64_t n, cudaStream_t st);
int so_build_tc_image(const so_surrogate *s, std::vector<unsigned char> &img, int *n_mma_out);
int so_launch_score_tc(const so_lattice *lat, const so_surrogate *s, const uint32_t *bits_dev, int64_t n, int32_t *hq_dev,
                       long long *hi, long long *lo, int32_t *nact, cudaStream_t st, int packed_out = 0);
int so_launch_objective(const so_surrogate *s, const long long *hi, const long long *lo, const int32_t *nact, int64_t n,
                        double *of_dev, cudaStream_t st);
int so_launch_eval_rows(const so_surrogate *s, int64_t n_rows, const uint8_t *X_dev, uint8_t *active_dev,
                        double *rate_dev, cudaStream_t st);
// kernel ids reported in so_anneal_stats.kernel_id (names: so_kernel_name)
enum { SO_K_NONE = 0, SO_K_TEAM = 1, SO_K_RESIDENT = 2, SO_K_GATE_INDEX = 3, SO_K_WINDOW = 4, SO_K_FLAT = 5, SO_K_GATED = 6,
       SO_K_ANALYTICAL = 7, SO_K_WINDOW_GATED = 8 };   // (the window kernel on the 24-bit packed state reports SO_K_WINDOW)
int so_launch_anneal(const so_chains *c, const SaParams &P, cudaStream_t st, int *kernel_id);
int so_launch_islands(uint32_t *bits_dev, int64_t n, int d, int N, int W, const int32_t *index_dev, int n_strucs, int n_mutate,
                      int32_t *scratch_dev, cudaStream_t st);
int so_launch_pack24(const int32_t *src_dev, void *dst_dev, int64_t n_sites, const int32_t *bias_dev, cudaStream_t st);
int so_launch_unpack24(const void *src_dev, int32_t *dst_dev, int64_t n_sites, const int32_t *bias_dev, cudaStream_t st);
int so_launch_anneal_window(const so_chains *c, const SaParams &P, cudaStream_t st);
bool so_window_gated_ok(const so_chains *c);
int so_launch_anneal_window_gated(const so_chains *c, const SaParams &P, cudaStream_t st, int build_records);
size_t so_window_smem_bytes(int Q, int W, int stages);
int so_launch_anneal_resident(const so_chains *c, const SaParams &P, cudaStream_t st);
bool so_resident_plan(const so_surrogate *s, int N, int W, int *warps, size_t *smem, int min_warps_per_sm = 8,
                      int *index_global = nullptr);
bool so_gate_index_ok(const so_surrogate *s, int N, int W);
bool so_team_ok(const so_surrogate *s, int N, int W, long long n_chains, int n_sm);
int so_launch_anneal_team(const so_chains *c, const SaParams &P, cudaStream_t st);
int so_launch_anneal_gate_index(const so_chains *c, const SaParams &P, cudaStream_t st);
int so_launch_apply_flips(const so_chains *c, int64_t n, const int64_t *chains_dev, const int32_t *sites_dev,
                          cudaStream_t st);
int so_lsc_max_sites(void);
int so_launch_lsc_analytical(const so_lattice *lat, const uint32_t *bits_dev, int64_t n, double *site_rates_dev,
                             double *struct_rate_dev, cudaStream_t st);
int so_launch_anneal_analytical(const so_chains *c, const SaParams &P, double *of_dev, cudaStream_t st);
int so_launch_best(const so_chains *c, int64_t chain_id0, void *record_dev, cudaStream_t st);
int so_launch_exchange(const so_chains *c, int rank, int world, int ladder, int parity, uint64_t seed, int64_t round,
                       double t_ref, const double *of_all, double *tscale_all, double *scratch,
                       unsigned long long *n_swapped, cudaStream_t st);

#ifdef __CUDACC__
// ---- device helpers ----------------------------------------------------------------------------------

// The pinned fp64 conversion of exact integer sums (oracle/surrogate.py real_from_sums): no fused ops.
__device__ __forceinline__ double so_real(double c, double b2, double sigma, double mu, long long hi, long long lo,
                                          int n) {
  double S = __dadd_rn(__dmul_rn(__ll2double_rn(hi), 1048576.0), __ll2double_rn(lo));
  double nn = (double)n;
  double t3 = __dadd_rn(__dmul_rn(c, S), __dmul_rn(b2, nn));
  return __dadd_rn(__dmul_rn(sigma, t3), __dmul_rn(mu, nn));
}

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ int wrap(int v, int d) {   // valid for -d <= v < 2d
  v += (v < 0) ? d : 0;
  v -= (v >= d) ? d : 0;
  return v;
}

// bit view of one chain's occupancy with an optional virtual flip
struct BitView {
  const uint32_t *w;
  int d;
  int flip;   // site index whose bit reads inverted, or -1
  __device__ __forceinline__ int site(int v) const { return ((w[v >> 5] >> (v & 31)) & 1) ^ (v == flip); }
  __device__ __forceinline__ int at(int i1, int i2) const { return site(wrap(i2, d) * d + wrap(i1, d)); }
};

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SO_FULL, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) { return __reduce_add_sync(SO_FULL, v); }

// ONE near-tie predicate for every SA kernel (exact_guard): the rule is "accept iff delta > thr" with thr = T ln u for
// T > 0 and u > 0 (no finite threshold for u = 0), thr = 0 for T <= 0; a decision is re-taken from the fp64 weights when
// the quantised delta is within tol of thr -- whatever its sign.
__device__ __forceinline__ bool so_near_tie(double delta, double T, float u, double tol) {
  if (T > 0.0) return u > 0.f && fabs(delta - T * log((double)u)) < tol;
  return fabs(delta) < tol;
}

// ---- tree gate ----------------------------------------------------------------------------------------
// Descriptor column c of row (r1,r2) is the site (r1+o1[c], r2+o2[c]).
// gate of row (r1,r2) before and after flipping site f in ONE walk: both walks share the path until a node tests the
// flipped site (if none does, the gate cannot change)
__device__ __forceinline__ void tree_gate_pair(const SurView &sv, const uint32_t *bits, int d, int r1, int r2, int f,
                                               int &g0, int &g1) {
  int4 nd = __ldg(sv.t_packed);
  while (!(nd.w & 1)) {
    int o1 = (int)(short)(nd.x & 0xffff), o2 = nd.x >> 16;
    int v = wrap(r2 + o2, d) * d + wrap(r1 + o1, d);
    int bit = (bits[v >> 5] >> (v & 31)) & 1;
    if (v == f) {                                   // the two structures part ways here
      int4 na = __ldg(sv.t_packed + (((nd.w >> (2 + bit)) & 1) ? nd.y : nd.z));
      int4 nb = __ldg(sv.t_packed + (((nd.w >> (3 - bit)) & 1) ? nd.y : nd.z));
      while (!(na.w & 1)) {
        int a1 = (int)(short)(na.x & 0xffff), a2 = na.x >> 16;
        int va = wrap(r2 + a2, d) * d + wrap(r1 + a1, d);
        int ba = (bits[va >> 5] >> (va & 31)) & 1;
        na = __ldg(sv.t_packed + (((na.w >> (2 + ba)) & 1) ? na.y : na.z));
      }
      while (!(nb.w & 1)) {
        int a1 = (int)(short)(nb.x & 0xffff), a2 = nb.x >> 16;
        int vb = wrap(r2 + a2, d) * d + wrap(r1 + a1, d);
        int bb = ((bits[vb >> 5] >> (vb & 31)) & 1) ^ (vb == f);
        nb = __ldg(sv.t_packed + (((nb.w >> (2 + bb)) & 1) ? nb.y : nb.z));
      }
      g0 = (na.w >> 1) & 1;
      g1 = (nb.w >> 1) & 1;
      return;
    }
    nd = __ldg(sv.t_packed + (((nd.w >> (2 + bit)) & 1) ? nd.y : nd.z));
  }
  g0 = g1 = (nd.w >> 1) & 1;
}

__device__ __forceinline__ int tree_gate(const SurView &sv, const BitView &b, int r1, int r2) {
  // one 16-byte record per node visit (the test "x <= thr" on a 0/1 input is two precomputed flag bits)
  int4 nd = __ldg(sv.t_packed);
  while (!(nd.w & 1)) {
    int o1 = (int)(short)(nd.x & 0xffff), o2 = nd.x >> 16;
    int bit = b.at(r1 + o1, r2 + o2);
    int go_left = (nd.w >> (2 + bit)) & 1;
    nd = __ldg(sv.t_packed + (go_left ? nd.y : nd.z));
  }
  return (nd.w >> 1) & 1;
}


// Exact fp64 delta from the bits and the ORIGINAL fp64 weights, for near-tie decisions (whole warp).
// Returns OF(x^e_f) - OF(x) as the difference of two window sums evaluated in the same lane order.
__device__ inline double so_exact_delta(const SurView &sv, const uint32_t *bits, int d, int f, int lane) {
  const int f2 = f / d, f1 = f - f2 * d;
  BitView b0{bits, d, -1}, b1{bits, d, f};
  double s_old = 0.0, s_new = 0.0;
  for (int t = lane; t < sv.K; t += 32) {
    int off = __ldg(sv.offs + t);
    int r1 = wrap(f1 - (int)(short)(off & 0xffff), d), r2 = wrap(f2 - (off >> 16), d);
    double y0 = sv.b2, y1 = sv.b2;
    for (int j = 0; j < sv.H; ++j) {
      double h0 = __ldg(sv.b1d + j), h1 = h0;
      for (int k = 0; k < sv.K; ++k) {
        int o = __ldg(sv.offs + k);
        int a1 = wrap(r1 + (int)(short)(o & 0xffff), d), a2 = wrap(r2 + (o >> 16), d);
        double w = __ldg(sv.W1d + (long long)k * sv.H + j);
        if (b0.site(a2 * d + a1)) h0 += w;
        if (b1.site(a2 * d + a1)) h1 += w;
      }
      double w2 = __ldg(sv.w2d + j);
      y0 += w2 * fmax(h0, 0.0);
      y1 += w2 * fmax(h1, 0.0);
    }
    int g0 = sv.gated ? tree_gate(sv, b0, r1, r2) : 1;
    int g1 = sv.gated ? tree_gate(sv, b1, r1, r2) : 1;
    if (g0) s_old += y0 * sv.y_scale + sv.y_mean;
    if (g1) s_new += y1 * sv.y_scale + sv.y_mean;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s_old += __shfl_xor_sync(SO_FULL, s_old, o);
    s_new += __shfl_xor_sync(SO_FULL, s_new, o);
  }
  return s_new - s_old;
}

#endif
