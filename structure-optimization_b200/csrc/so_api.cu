// so_api.cu -- the extern "C" boundary (include/so_b200.h): handles, host<->device staging, error mapping.
// Everything computational is a CUDA kernel in so_desc.cu / so_sa.cu; there is no CPU fallback.
#include "so_internal.cuh"
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <new>
#include <vector>
#include <nvtx3/nvToolsExt.h>

static thread_local std::string g_err;

void so_set_error(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

namespace {

const double A_LAT = 3.9239;            // OML/dynamic_cat.py:49
const double VACUUM = 15.0;
const int RING[6][2] = {{1, 0}, {0, 1}, {-1, 1}, {-1, 0}, {0, -1}, {1, -1}};
const int UP_34[3][2] = {{0, 0}, {-1, 1}, {0, 1}};
const int DN_43[3][2] = {{0, -1}, {1, -1}, {0, 0}};
const int DN_32[3][2] = {{-1, 0}, {0, 0}, {-1, 1}};
const int UP_23[3][2] = {{1, -1}, {0, 0}, {1, 0}};
const int UP_12[3][2] = {{-1, 0}, {0, -1}, {0, 0}};
const int DN_21[3][2] = {{1, 0}, {0, 1}, {0, 0}};
const double SHIFT[5][2] = {{-1.0 / 3, 2.0 / 3}, {0, 0}, {1.0 / 3, 1.0 / 3}, {-1.0 / 3, 2.0 / 3}, {0, 0}};

inline int wrapi(int v, int d) { v %= d; return v < 0 ? v + d : v; }

int ceil_log2(double x) {   // exact: smallest e with 2^e >= x
  int e;
  double m = frexp(x, &e);  // x = m * 2^e, 0.5 <= m < 1
  return (m == 0.5) ? e - 1 : e;
}

}  // namespace

// ---- helpers with C++ linkage
template <class T>
static int upload(T **dst, const T *src, size_t n) {
  *dst = nullptr;
  cudaError_t e = cudaMalloc((void **)dst, std::max<size_t>(n, 1) * sizeof(T));
  if (e != cudaSuccess) { so_set_error("cudaMalloc failed: %s", cudaGetErrorString(e)); return SO_ENOMEM; }
  if (n) SO_CUDA(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return SO_OK;
}

extern "C" {

int so_version(void) { return 100; }
const char *so_last_error(void) { return g_err.c_str(); }

int so_device_count(int *n_out) {
  SO_REQUIRE(n_out, "n_out is NULL");
  SO_CUDA(cudaGetDeviceCount(n_out));
  return SO_OK;
}

// ---- lattice -----------------------------------------------------------------------------------------
int so_lattice_create(int d, int kind, int device, so_lattice **out) {
  SO_REQUIRE(out, "out is NULL");
  SO_REQUIRE(d >= 4 && d <= 1024, "lattice edge d=%d out of range [4,1024] (closed-form typing needs d >= 4)", d);
  SO_REQUIRE(kind == SO_KIND_LSC || kind == SO_KIND_NH3, "unknown catalyst kind %d", kind);
  SO_CUDA(cudaSetDevice(device));
  {
    // The SA state is read in 560-byte runs that start on 16-byte boundaries: with the default 64-byte L2 fetch
    // granule every run drags in ~48 bytes it does not need.  Ask for 32-byte fills (a hint the platform may
    // clamp); SO_L2_FETCH_GRANULARITY = 0 leaves the default, 32 / 64 / 128 select a value.
    const char *e = getenv("SO_L2_FETCH_GRANULARITY");
    const int want = e ? atoi(e) : 32;
    if (want > 0 && cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)want) != cudaSuccess) cudaGetLastError();
  }
  so_lattice *L = new (std::nothrow) so_lattice();
  if (!L) { so_set_error("out of host memory"); return SO_ENOMEM; }
  L->d = d; L->N = d * d; L->W = (L->N + 31) / 32; L->kind = kind; L->device = device;
  L->h5_ptr_dev = L->h5_idx_dev = nullptr;
  const int N = L->N;
  const double dnn = A_LAT / sqrt(2.0), dz = A_LAT / sqrt(3.0);
  L->cell[0] = d * dnn; L->cell[1] = 0.0; L->cell[2] = d * dnn * 0.5; L->cell[3] = d * dnn * sqrt(0.75);
  L->pos.resize((size_t)5 * N * 3);
  for (int layer = 0; layer < 5; ++layer)
    for (int i2 = 0; i2 < d; ++i2)
      for (int i1 = 0; i1 < d; ++i1) {
        double u = i1 + SHIFT[layer][0], v = i2 + SHIFT[layer][1];
        size_t a = (size_t)layer * N + (size_t)i2 * d + i1;
        L->pos[3 * a + 0] = dnn * (u + 0.5 * v);
        L->pos[3 * a + 1] = dnn * sqrt(0.75) * v;
        L->pos[3 * a + 2] = VACUUM + layer * dz;
      }
  // neighbour CSR: ring, then the layer above, then the layer below
  L->rowptr.reserve((size_t)5 * N + 1);
  L->col.reserve((size_t)54 * N);
  L->rowptr.push_back(0);
  auto node = [&](int layer, int i1, int i2) { return layer * N + wrapi(i2, d) * d + wrapi(i1, d); };
  for (int layer = 0; layer < 5; ++layer)
    for (int i2 = 0; i2 < d; ++i2)
      for (int i1 = 0; i1 < d; ++i1) {
        for (auto &r : RING) L->col.push_back(node(layer, i1 + r[0], i2 + r[1]));
        const int(*up)[2] = layer == 0 ? UP_34 : layer == 1 ? UP_12 : layer == 2 ? UP_23 : layer == 3 ? UP_34 : nullptr;
        const int(*dn)[2] = layer == 1 ? DN_43 : layer == 2 ? DN_21 : layer == 3 ? DN_32 : layer == 4 ? DN_43 : nullptr;
        if (up) for (int k = 0; k < 3; ++k) L->col.push_back(node(layer + 1, i1 + up[k][0], i2 + up[k][1]));
        if (dn) for (int k = 0; k < 3; ++k) L->col.push_back(node(layer - 1, i1 + dn[k][0], i2 + dn[k][1]));
        L->rowptr.push_back((int32_t)L->col.size());
      }
  // NH3 type 9: layer-2 atoms within 2*2.77 A of each layer-3 site on RAW xy positions (OML/NH3_cat.py:367-381).
  // Raw positions are translation invariant away from the cell edge, so only a 5x5 index neighbourhood can qualify.
  std::vector<int32_t> ptr(N + 1, 0), idx;
  for (int i2 = 0; i2 < d; ++i2)
    for (int i1 = 0; i1 < d; ++i1) {
      size_t a = (size_t)3 * N + (size_t)i2 * d + i1;
      for (int k2 = std::max(0, i2 - 3); k2 <= std::min(d - 1, i2 + 3); ++k2)
        for (int k1 = std::max(0, i1 - 3); k1 <= std::min(d - 1, i1 + 3); ++k1) {
          size_t b = (size_t)2 * N + (size_t)k2 * d + k1;
          double dx = L->pos[3 * a] - L->pos[3 * b], dy = L->pos[3 * a + 1] - L->pos[3 * b + 1];
          if (sqrt(dx * dx + dy * dy) < 2.77 * 2) idx.push_back(k2 * d + k1);
        }
      ptr[i2 * d + i1 + 1] = (int32_t)idx.size();
    }
  if (cudaMalloc(&L->h5_ptr_dev, ptr.size() * 4) != cudaSuccess ||
      cudaMalloc(&L->h5_idx_dev, std::max<size_t>(idx.size(), 1) * 4) != cudaSuccess) {
    so_set_error("cudaMalloc failed for lattice tables: %s", cudaGetErrorString(cudaGetLastError()));
    so_lattice_destroy(L);
    return SO_ENOMEM;
  }
  SO_CUDA(cudaMemcpy(L->h5_ptr_dev, ptr.data(), ptr.size() * 4, cudaMemcpyHostToDevice));
  if (!idx.empty()) SO_CUDA(cudaMemcpy(L->h5_idx_dev, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice));
  *out = L;
  return SO_OK;
}

int so_lattice_destroy(so_lattice *L) {
  if (!L) return SO_OK;
  cudaSetDevice(L->device);
  if (L->h5_ptr_dev) cudaFree(L->h5_ptr_dev);
  if (L->h5_idx_dev) cudaFree(L->h5_idx_dev);
  delete L;
  return SO_OK;
}

int so_lattice_info(const so_lattice *L, int *d, int *n_sites, int *n_nodes, int *nnz, int *words) {
  SO_REQUIRE(L, "lattice is NULL");
  if (d) *d = L->d;
  if (n_sites) *n_sites = L->N;
  if (n_nodes) *n_nodes = 5 * L->N;
  if (nnz) *nnz = (int)L->col.size();
  if (words) *words = L->W;
  return SO_OK;
}

int so_lattice_csr(const so_lattice *L, int32_t *rowptr, int32_t *col) {
  SO_REQUIRE(L && rowptr && col, "NULL argument");
  memcpy(rowptr, L->rowptr.data(), L->rowptr.size() * 4);
  memcpy(col, L->col.data(), L->col.size() * 4);
  return SO_OK;
}

int so_lattice_positions(const so_lattice *L, double *pos, double *cell) {
  SO_REQUIRE(L, "lattice is NULL");
  if (pos) memcpy(pos, L->pos.data(), L->pos.size() * 8);
  if (cell) memcpy(cell, L->cell, 4 * 8);
  return SO_OK;
}

// ---- chains --------------------------------------------------------------------------------------------
int so_chains_create(so_lattice *L, int64_t n_chains, so_chains **out) {
  SO_REQUIRE(L && out, "NULL argument");
  SO_REQUIRE(n_chains >= 1 && n_chains <= (1LL << 31), "n_chains=%lld out of range", (long long)n_chains);
  SO_CUDA(cudaSetDevice(L->device));
  so_chains *c = new (std::nothrow) so_chains();
  if (!c) { so_set_error("out of host memory"); return SO_ENOMEM; }
  memset(c, 0, sizeof(*c));
  c->lat = L; c->device = L->device; c->n_chains = n_chains;
  cudaError_t e = cudaMalloc(&c->bits_dev, (size_t)n_chains * L->W * 4);
  if (e == cudaSuccess) e = cudaMalloc(&c->tscale_dev, (size_t)n_chains * 8);
  if (e == cudaSuccess) e = cudaMalloc(&c->counters_dev, 8 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
  if (e != cudaSuccess) {
    so_set_error("chain allocation failed: %s", cudaGetErrorString(e));
    so_chains_destroy(c);
    return SO_ENOMEM;
  }
  // all atoms present (variable_occs = [1]*N, OML/LSC_cat.py:50), tail bits of the last word zero
  std::vector<uint32_t> ones(L->W, 0xffffffffu);
  if (L->N & 31) ones[L->W - 1] = (1u << (L->N & 31)) - 1u;
  std::vector<uint32_t> all((size_t)std::min<int64_t>(n_chains, 4096) * L->W);
  for (size_t i = 0; i < all.size(); ++i) all[i] = ones[i % L->W];
  for (int64_t c0 = 0; c0 < n_chains; c0 += 4096) {
    int64_t nc = std::min<int64_t>(4096, n_chains - c0);
    SO_CUDA(cudaMemcpy(c->bits_dev + c0 * L->W, all.data(), (size_t)nc * L->W * 4, cudaMemcpyHostToDevice));
  }
  std::vector<double> one((size_t)n_chains, 1.0);
  SO_CUDA(cudaMemcpy(c->tscale_dev, one.data(), (size_t)n_chains * 8, cudaMemcpyHostToDevice));
  SO_CUDA(cudaMemset(c->counters_dev, 0, 8 * sizeof(unsigned long long)));
  *out = c;
  return SO_OK;
}

int so_chains_destroy(so_chains *c) {
  if (!c) return SO_OK;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  cudaFree(c->bits_dev); cudaFree(c->hq_dev); cudaFree(c->hi_dev); cudaFree(c->lo_dev); cudaFree(c->nact_dev);
  cudaFree(c->tscale_dev); cudaFree(c->counters_dev); cudaFree(c->xch_scratch); cudaFree(c->gate_scratch); cudaFree(c->temps_dev);
  cudaFree(c->type_hist_dev);
  cudaFree(c->tmp_hq_dev);
  cudaFree(c->gate_meta_dev);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return SO_OK;
}

static int check_range(const so_chains *c, int64_t chain0, int64_t n) {
  SO_REQUIRE(c, "chains handle is NULL");
  SO_REQUIRE(chain0 >= 0 && n >= 0 && chain0 + n <= c->n_chains, "chain range [%lld, %lld) outside [0, %lld)",
             (long long)chain0, (long long)(chain0 + n), (long long)c->n_chains);
  return SO_OK;
}

// after the bits of a range changed: keep the surrogate state consistent
// NVTX ranges (header-only nvtx3: a no-op unless a tool is attached) around the phases a timeline should show: state
// build, sweep, the kernels that feed a collective.  The reference's only timing hook is time.time() around its loop
// (OML/optimize_SA.py:55,135-136).
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

// tracked type histograms (so_track_types) follow every change of the bits: recomputed here, carried by the SA kernels
static int refresh_types(so_chains *c, int64_t chain0, int64_t n) {
  if (!c->type_hist_dev || n == 0) return SO_OK;
  const so_lattice *L = c->lat;
  return so_launch_descriptors(L, c->bits_dev + chain0 * L->W, n, nullptr, nullptr, nullptr,
                               c->type_hist_dev + c->n_chains * 8 + chain0 * 3, nullptr, c->type_hist_dev + chain0 * 8,
                               c->stream);
}

static int rescore(so_chains *c, int64_t chain0, int64_t n) {
  c->meta_valid = 0;                               // the bits changed: the gated window kernel rebuilds its row records
  SO_TRY(refresh_types(c, chain0, n));
  if (!c->sur || n == 0) return SO_OK;
  NvtxRange nvtx_range("state build (first-layer state + objective sums from the bits)");
  const so_lattice *L = c->lat;
  if (c->sur->compact) {
    // the tensor-core scorer packs its rows to 24 bits / 64 bytes per site in the epilogue (SO_SCORE_TC=0 / SO_SCORE_PACK=0 keep
    // the two-pass build below: both give the same bytes, tests compare them)
    {
      const char *e1 = getenv("SO_SCORE_TC"), *e2 = getenv("SO_SCORE_PACK");
      if (c->sur->tc_ok && !(e1 && e1[0] == '0') && !(e2 && e2[0] == '0'))
        return so_launch_score_tc(L, c->sur, c->bits_dev + chain0 * L->W, n,
                                  (int32_t *)((char *)c->hq_dev + (size_t)chain0 * L->N * 64), c->hi_dev + chain0,
                                  c->lo_dev + chain0, c->nact_dev + chain0, c->stream, 1);
    }
    // otherwise the scorers write int32 [site][20]; slabs of it are packed to 24 bits / 64 bytes per site (so_sa_window.cu)
    const int64_t slab = std::max<int64_t>(1, std::min<int64_t>(n, (1LL << 30) / ((int64_t)L->N * 80)));
    if ((int64_t)c->tmp_hq_chains < slab) {
      cudaFree(c->tmp_hq_dev);
      c->tmp_hq_dev = nullptr; c->tmp_hq_chains = 0;
      if (cudaMalloc((void **)&c->tmp_hq_dev, (size_t)slab * L->N * 80) != cudaSuccess) {
        cudaGetLastError();
        so_set_error("cudaMalloc failed for the state-build staging slab");
        return SO_ENOMEM;
      }
      c->tmp_hq_chains = (size_t)slab;
    }
    for (int64_t c0 = 0; c0 < n; c0 += slab) {
      const int64_t nc = std::min(slab, n - c0);
      SO_TRY(so_launch_score(L, c->sur, c->bits_dev + (chain0 + c0) * L->W, nc, c->tmp_hq_dev, c->hi_dev + chain0 + c0,
                             c->lo_dev + chain0 + c0, c->nact_dev + chain0 + c0, c->stream));
      SO_TRY(so_launch_pack24(c->tmp_hq_dev, (char *)c->hq_dev + (size_t)(chain0 + c0) * L->N * 64, nc * L->N, c->sur->bias24_dev,
                              c->stream));
    }
    return SO_OK;
  }
  return so_launch_score(L, c->sur, c->bits_dev + chain0 * L->W, n,
                         c->hq_dev + chain0 * (long long)L->N * c->sur->Hp, c->hi_dev + chain0, c->lo_dev + chain0,
                         c->nact_dev + chain0, c->stream);
}

int so_set_occupancy(so_chains *c, int64_t chain0, int64_t n, const uint8_t *occ) {
  SO_TRY(check_range(c, chain0, n));
  SO_REQUIRE(occ || n == 0, "occ is NULL");
  const so_lattice *L = c->lat;
  SO_CUDA(cudaSetDevice(L->device));
  const int64_t slab = std::max<int64_t>(1, (256LL << 20) / L->N);
  DevBuf stage;
  SO_TRY(stage.alloc((size_t)std::min(slab, std::max<int64_t>(n, 1)) * L->N));
  SO_CUDA(cudaMemsetAsync(c->counters_dev + 3, 0, sizeof(unsigned long long), c->stream));
  for (int64_t c0 = 0; c0 < n; c0 += slab) {
    int64_t nc = std::min(slab, n - c0);
    SO_CUDA(cudaMemcpyAsync(stage.p, occ + c0 * L->N, (size_t)nc * L->N, cudaMemcpyHostToDevice, c->stream));
    SO_TRY(so_launch_pack(stage.as<uint8_t>(), c->bits_dev + (chain0 + c0) * L->W, nc, L->N, L->W, c->counters_dev + 3,
                          c->stream));
    SO_CUDA(cudaStreamSynchronize(c->stream));
  }
  unsigned long long bad = 0;
  SO_CUDA(cudaMemcpy(&bad, c->counters_dev + 3, sizeof(bad), cudaMemcpyDeviceToHost));
  SO_TRY(rescore(c, chain0, n));
  SO_CUDA(cudaStreamSynchronize(c->stream));
  SO_REQUIRE(!bad, "Invalid occupancy");   // OML/dynamic_cat.py:185 raises NameError('Invalid occupancy.')
  return SO_OK;
}

int so_get_occupancy(so_chains *c, int64_t chain0, int64_t n, uint8_t *occ) {
  SO_TRY(check_range(c, chain0, n));
  SO_REQUIRE(occ || n == 0, "occ is NULL");
  const so_lattice *L = c->lat;
  SO_CUDA(cudaSetDevice(L->device));
  const int64_t slab = std::max<int64_t>(1, (256LL << 20) / L->N);
  DevBuf stage;
  SO_TRY(stage.alloc((size_t)std::min(slab, std::max<int64_t>(n, 1)) * L->N));
  for (int64_t c0 = 0; c0 < n; c0 += slab) {
    int64_t nc = std::min(slab, n - c0);
    SO_TRY(so_launch_unpack(c->bits_dev + (chain0 + c0) * L->W, stage.as<uint8_t>(), nc, L->N, L->W, c->stream));
    SO_CUDA(cudaMemcpyAsync(occ + c0 * L->N, stage.p, (size_t)nc * L->N, cudaMemcpyDeviceToHost, c->stream));
    SO_CUDA(cudaStreamSynchronize(c->stream));
  }
  return SO_OK;
}

int so_set_occupancy_bits(so_chains *c, int64_t chain0, int64_t n, const uint32_t *bits) {
  SO_TRY(check_range(c, chain0, n));
  SO_REQUIRE(bits || n == 0, "bits is NULL");
  const so_lattice *L = c->lat;
  SO_CUDA(cudaSetDevice(L->device));
  if ((L->N & 31) && n > 0) {   // bits beyond site N-1 must be zero
    uint32_t tail = ~((1u << (L->N & 31)) - 1u);
    for (int64_t i = 0; i < n; ++i)
      SO_REQUIRE((bits[i * L->W + L->W - 1] & tail) == 0, "Invalid occupancy (bits set beyond the last site)");
  }
  SO_CUDA(cudaMemcpyAsync(c->bits_dev + chain0 * L->W, bits, (size_t)n * L->W * 4, cudaMemcpyHostToDevice, c->stream));
  SO_TRY(rescore(c, chain0, n));
  SO_CUDA(cudaStreamSynchronize(c->stream));
  return SO_OK;
}

int so_get_occupancy_bits(so_chains *c, int64_t chain0, int64_t n, uint32_t *bits) {
  SO_TRY(check_range(c, chain0, n));
  SO_REQUIRE(bits || n == 0, "bits is NULL");
  const so_lattice *L = c->lat;
  SO_CUDA(cudaSetDevice(L->device));
  SO_CUDA(cudaMemcpyAsync(bits, c->bits_dev + chain0 * L->W, (size_t)n * L->W * 4, cudaMemcpyDeviceToHost, c->stream));
  SO_CUDA(cudaStreamSynchronize(c->stream));
  return SO_OK;
}

int so_randomize(so_chains *c, uint64_t seed, double coverage) {
  SO_REQUIRE(c, "chains handle is NULL");
  SO_REQUIRE(coverage <= 1.0, "coverage %g > 1", coverage);
  const so_lattice *L = c->lat;
  SO_CUDA(cudaSetDevice(L->device));
  SO_TRY(so_launch_randomize(c->bits_dev, c->n_chains, L->N, L->W, seed, coverage, c->stream));
  SO_TRY(rescore(c, 0, c->n_chains));
  SO_CUDA(cudaStreamSynchronize(c->stream));
  return SO_OK;
}

int so_island_structures(so_chains *c, int64_t chain0, int64_t n, const int32_t *index, int n_strucs, int n_mutate) {
  SO_TRY(check_range(c, chain0, n));
  SO_REQUIRE(index || n == 0, "index is NULL");
  SO_REQUIRE(n_strucs >= 1 && n_mutate >= 0, "n_strucs must be positive and n_mutate non-negative");
  for (int64_t k = 0; k < n; ++k) SO_REQUIRE(index[k] >= 0, "structure index %lld is negative (np.random.seed needs >= 0)", (long long)k);
  if (n == 0) return SO_OK;
  const so_lattice *L = c->lat;
  SO_CUDA(cudaSetDevice(L->device));
  const int64_t slab = std::max<int64_t>(1, (1LL << 30) / ((int64_t)L->N * 4));      // <= 1 GiB of permutation scratch
  DevBuf d_idx, d_perm;
  SO_TRY(d_idx.alloc((size_t)n * 4));
  SO_TRY(d_perm.alloc((size_t)std::min(slab, n) * L->N * 4));
  SO_CUDA(cudaMemcpyAsync(d_idx.p, index, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  for (int64_t k0 = 0; k0 < n; k0 += slab) {
    const int64_t nk = std::min(slab, n - k0);
    SO_TRY(so_launch_islands(c->bits_dev + (chain0 + k0) * L->W, nk, L->d, L->N, L->W, d_idx.as<int32_t>() + k0, n_strucs,
                             n_mutate, d_perm.as<int32_t>(), c->stream));
  }
  SO_TRY(rescore(c, chain0, n));
  SO_CUDA(cudaStreamSynchronize(c->stream));
  return SO_OK;
}

int so_flip(so_chains *c, int64_t n, const int64_t *chains, const int32_t *sites) {
  SO_REQUIRE(c, "chains handle is NULL");
  SO_REQUIRE(n >= 0 && (n == 0 || (chains && sites)), "bad flip list");
  const so_lattice *L = c->lat;
  for (int64_t i = 0; i < n; ++i) {
    SO_REQUIRE(chains[i] >= 0 && chains[i] < c->n_chains, "flip %lld: chain %lld out of range", (long long)i,
               (long long)chains[i]);
    SO_REQUIRE(sites[i] >= 0 && sites[i] < L->N, "flip %lld: site %d out of range", (long long)i, sites[i]);
  }
  if (n == 0) return SO_OK;
  SO_CUDA(cudaSetDevice(L->device));
  DevBuf dc, ds;
  SO_TRY(dc.alloc((size_t)n * 8));
  SO_TRY(ds.alloc((size_t)n * 4));
  SO_CUDA(cudaMemcpyAsync(dc.p, chains, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
  SO_CUDA(cudaMemcpyAsync(ds.p, sites, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  SO_TRY(so_launch_apply_flips(c, n, dc.as<int64_t>(), ds.as<int32_t>(), c->stream));
  if (c->sur) {
    std::vector<int64_t> uniq(chains, chains + n);
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    if ((int64_t)uniq.size() > 64) {
      SO_TRY(rescore(c, 0, c->n_chains));
    } else {
      for (int64_t ch : uniq) SO_TRY(rescore(c, ch, 1));
    }
  }
  SO_CUDA(cudaStreamSynchronize(c->stream));
  return SO_OK;
}

// ---- descriptors ----------------------------------------------------------------------------------------
int so_descriptors(so_chains *c, int64_t chain0, int64_t n, uint8_t *c6, uint8_t *lsc_type, uint8_t *nh3_type,
                   int32_t *lsc_hist, int32_t *nh3_hist, int32_t *nh3_exp) {
  SO_TRY(check_range(c, chain0, n));
  const so_lattice *L = c->lat;
  SO_CUDA(cudaSetDevice(L->device));
  const int N = L->N;
  // staging slab: 256 MB of the outputs actually requested (histograms alone are 124 B per structure: one launch)
  const int64_t out_bytes = (c6 ? N : 0) + (lsc_type ? N : 0) + (nh3_type ? 5LL * N : 0) + 124;
  const int64_t slab = std::max<int64_t>(1, std::min<int64_t>(1LL << 22, (256LL << 20) / out_bytes));
  DevBuf b_c6, b_lsc, b_nh3, b_lh, b_nh, b_ne;
  int64_t cap = std::min(slab, std::max<int64_t>(n, 1));
  if (c6) SO_TRY(b_c6.alloc((size_t)cap * N));
  if (lsc_type) SO_TRY(b_lsc.alloc((size_t)cap * N));
  if (nh3_type) SO_TRY(b_nh3.alloc((size_t)cap * 5 * N));
  if (lsc_hist) SO_TRY(b_lh.alloc((size_t)cap * 3 * 4));
  if (nh3_hist) SO_TRY(b_nh.alloc((size_t)cap * 20 * 4));
  if (nh3_exp) SO_TRY(b_ne.alloc((size_t)cap * 8 * 4));
  for (int64_t c0 = 0; c0 < n; c0 += slab) {
    int64_t nc = std::min(slab, n - c0);
    if (nh3_type) SO_CUDA(cudaMemsetAsync(b_nh3.p, 0, (size_t)nc * 5 * N, c->stream));   // layer 0 = None
    SO_TRY(so_launch_descriptors(L, c->bits_dev + (chain0 + c0) * L->W, nc, b_c6.as<uint8_t>(), b_lsc.as<uint8_t>(),
                                 b_nh3.as<uint8_t>(), b_lh.as<int32_t>(), b_nh.as<int32_t>(), b_ne.as<int32_t>(),
                                 c->stream));
    if (c6) SO_CUDA(cudaMemcpyAsync(c6 + c0 * N, b_c6.p, (size_t)nc * N, cudaMemcpyDeviceToHost, c->stream));
    if (lsc_type) SO_CUDA(cudaMemcpyAsync(lsc_type + c0 * N, b_lsc.p, (size_t)nc * N, cudaMemcpyDeviceToHost, c->stream));
    if (nh3_type)
      SO_CUDA(cudaMemcpyAsync(nh3_type + c0 * 5 * N, b_nh3.p, (size_t)nc * 5 * N, cudaMemcpyDeviceToHost, c->stream));
    if (lsc_hist) SO_CUDA(cudaMemcpyAsync(lsc_hist + c0 * 3, b_lh.p, (size_t)nc * 12, cudaMemcpyDeviceToHost, c->stream));
    if (nh3_hist) SO_CUDA(cudaMemcpyAsync(nh3_hist + c0 * 20, b_nh.p, (size_t)nc * 80, cudaMemcpyDeviceToHost, c->stream));
    if (nh3_exp) SO_CUDA(cudaMemcpyAsync(nh3_exp + c0 * 8, b_ne.p, (size_t)nc * 32, cudaMemcpyDeviceToHost, c->stream));
    SO_CUDA(cudaStreamSynchronize(c->stream));
  }
  return SO_OK;
}

int so_flip_delta(so_chains *c, int64_t chain0, int64_t n, const int32_t *sites, int32_t *d_lsc, int32_t *d_nh3_exp) {
  SO_TRY(check_range(c, chain0, n));
  SO_REQUIRE(sites || n == 0, "sites is NULL");
  const so_lattice *L = c->lat;
  for (int64_t i = 0; i < n; ++i) SO_REQUIRE(sites[i] >= 0 && sites[i] < L->N, "site %d out of range", sites[i]);
  if (n == 0) return SO_OK;
  SO_CUDA(cudaSetDevice(L->device));
  DevBuf ds, dl, dn;
  SO_TRY(ds.alloc((size_t)n * 4));
  if (d_lsc) SO_TRY(dl.alloc((size_t)n * 12));
  if (d_nh3_exp) SO_TRY(dn.alloc((size_t)n * 32));
  SO_CUDA(cudaMemcpyAsync(ds.p, sites, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  SO_TRY(so_launch_flip_delta(L, c->bits_dev + chain0 * L->W, n, ds.as<int32_t>(), dl.as<int32_t>(), dn.as<int32_t>(),
                              c->stream));
  if (d_lsc) SO_CUDA(cudaMemcpyAsync(d_lsc, dl.p, (size_t)n * 12, cudaMemcpyDeviceToHost, c->stream));
  if (d_nh3_exp) SO_CUDA(cudaMemcpyAsync(d_nh3_exp, dn.p, (size_t)n * 32, cudaMemcpyDeviceToHost, c->stream));
  SO_CUDA(cudaStreamSynchronize(c->stream));
  return SO_OK;
}

// ---- surrogate ------------------------------------------------------------------------------------------
static int surrogate_create_impl(so_lattice *L, int K, int H, const int32_t *offsets, const double *W1, const double *b1,
                                 const double *W2, double b2, double y_mean, double y_scale, int n_tree_nodes,
                                 const int32_t *feature, const double *thr, const int32_t *left, const int32_t *right,
                                 const uint8_t *leaf_active, int state_bits, so_surrogate **out) {
  SO_REQUIRE(L && offsets && W1 && b1 && W2 && out, "NULL argument");
  SO_REQUIRE(state_bits == 30 || state_bits == 22, "state_bits must be 30 (int32 state) or 22 (24-bit packed state)");
  SO_REQUIRE(K >= 1 && H >= 1 && H <= 32, "surrogate shape K=%d H=%d not supported (1 <= H <= 32)", K, H);
  SO_REQUIRE(n_tree_nodes >= 0, "negative tree size");
  SO_REQUIRE(n_tree_nodes == 0 || (feature && thr && left && right && leaf_active), "tree arrays missing");
  const int d = L->d;
  // rows r = f - offset_k must be distinct for distinct k (one input per row changes per flip)
  {
    std::vector<int> seen((size_t)L->N, 0);
    for (int k = 0; k < K; ++k) {
      int o1 = offsets[2 * k], o2 = offsets[2 * k + 1];
      SO_REQUIRE(o1 > -d && o1 < d && o2 > -d && o2 < d, "offset %d = (%d,%d) outside (-d, d)", k, o1, o2);
      int v = wrapi(o2, d) * d + wrapi(o1, d);
      SO_REQUIRE(!seen[v], "descriptor offsets %d aliases another column on a %dx%d lattice", k, d, d);
      seen[v] = 1;
    }
  }
  for (int i = 0; i < n_tree_nodes; ++i) {
    SO_REQUIRE(feature[i] < K, "tree node %d tests feature %d >= K", i, feature[i]);
    if (feature[i] >= 0)
      SO_REQUIRE(left[i] >= 0 && left[i] < n_tree_nodes && right[i] >= 0 && right[i] < n_tree_nodes,
                 "tree node %d has a child out of range", i);
  }
  SO_CUDA(cudaSetDevice(L->device));
  so_surrogate *s = new (std::nothrow) so_surrogate();
  if (!s) { so_set_error("out of host memory"); return SO_ENOMEM; }
  s->lat = L; s->device = L->device; s->K = K; s->H = H; s->Hp = (H + 3) / 4 * 4; s->HP4 = s->Hp / 4;
  s->b2 = b2; s->y_mean = y_mean; s->y_scale = y_scale; s->gated = n_tree_nodes > 0; s->n_tree = n_tree_nodes;
  s->offs_dev = nullptr; s->W1q_dev = s->b1q_dev = s->w2q_dev = nullptr; s->W1d_dev = s->b1d_dev = s->w2d_dev = nullptr;
  s->t_feature_dev = s->t_left_dev = s->t_right_dev = nullptr; s->t_thr_dev = nullptr; s->t_active_dev = nullptr;
  s->t_packed_dev = nullptr;
  s->wtree_dev = nullptr;
  // fixed-point grids (identical to oracle/surrogate.py quantize)
  double bound = 0.0, wmax = 0.0;
  for (int j = 0; j < H; ++j) {
    double sum = fabs(b1[j]);
    for (int k = 0; k < K; ++k) sum += fabs(W1[(size_t)k * H + j]);
    bound = std::max(bound, sum);
    wmax = std::max(wmax, fabs(W2[j]));
  }
  bound = std::max(bound, ldexp(1.0, -100));
  wmax = std::max(wmax, ldexp(1.0, -100));
  s->e_h = ceil_log2(bound) - 30;               // |hq| <= 2^30 (+ rounding slack), oracle: quantize()
  s->compact = state_bits == 22;
  s->Wst24_dev = nullptr; s->bias24_dev = nullptr;
  if (s->compact) {
    // packed state: 24 bits hold hq - (smallest reachable hq) of every unit, i.e. the unit's L1 range sum_k |W1q_kj|:
    // the grid is the finest power of two with max_j sum_k |rint(W1_kj / grid)| < 2^24 (oracle: quantize(bits='packed24'))
    double l1max = ldexp(1.0, -100);
    for (int j = 0; j < H; ++j) {
      double sum = 0.0;
      for (int k = 0; k < K; ++k) sum += fabs(W1[(size_t)k * H + j]);
      l1max = std::max(l1max, sum);
    }
    s->e_h = ceil_log2(l1max) - 24;
    for (;;) {
      long long worst = 0;
      for (int j = 0; j < H; ++j) {
        long long sum = 0;
        for (int k = 0; k < K; ++k) sum += llabs((long long)nearbyint(ldexp(W1[(size_t)k * H + j], -s->e_h)));
        worst = std::max(worst, sum);
      }
      if (worst < (1LL << 24)) break;
      ++s->e_h;
    }
  }
  int w_bits = 31 - ceil_log2((double)s->Hp);
  s->e_w = ceil_log2(wmax) - w_bits;
  s->c = ldexp(1.0, s->e_h + s->e_w);
  const int Hp = s->Hp;
  s->W1q_host.assign((size_t)K * Hp, 0); s->b1q_host.assign(Hp, 0); s->w2q_host.assign(Hp, 0); s->offs_host.resize(K);
  for (int k = 0; k < K; ++k) {
    for (int j = 0; j < H; ++j) s->W1q_host[(size_t)k * Hp + j] = (int32_t)nearbyint(ldexp(W1[(size_t)k * H + j], -s->e_h));
    s->offs_host[k] = (int32_t)(((uint32_t)offsets[2 * k + 1] << 16) | ((uint32_t)offsets[2 * k] & 0xffffu));
  }
  for (int j = 0; j < H; ++j) {
    s->b1q_host[j] = (int32_t)nearbyint(ldexp(b1[j], -s->e_h));
    s->w2q_host[j] = (int32_t)nearbyint(ldexp(W2[j], -s->e_w));
  }
  // rectangular-window descriptors get the pipelined kernel: rows in memory order p = p2*L + p1 <-> offset (b1-p1, b2-p2)
  s->Wst_dev = nullptr;
  std::vector<int32_t> Wst((size_t)K * Hp, 0);
  {
    int a1 = offsets[0], b1o = offsets[0], a2 = offsets[1], b2o = offsets[1];
    for (int k = 0; k < K; ++k) {
      a1 = std::min(a1, offsets[2 * k]); b1o = std::max(b1o, offsets[2 * k]);
      a2 = std::min(a2, offsets[2 * k + 1]); b2o = std::max(b2o, offsets[2 * k + 1]);
    }
    int Lw = b1o - a1 + 1, nseg = b2o - a2 + 1;
    s->win_ok = (Lw * nseg == K) && Lw <= d && nseg <= d && nseg <= 32;   // K distinct offsets in the box => full box
    s->win_a1 = a1; s->win_b1 = b1o; s->win_a2 = a2; s->win_b2 = b2o;
    if (s->win_ok)
      for (int k = 0; k < K; ++k) {
        int p = (b2o - offsets[2 * k + 1]) * Lw + (b1o - offsets[2 * k]);
        memcpy(&Wst[(size_t)p * Hp], &s->W1q_host[(size_t)k * Hp], (size_t)Hp * 4);
      }
  }
  int rc = SO_OK;
  s->tc_image_dev = nullptr; s->tc_ok = 0;
  {
    std::vector<unsigned char> img;
    int n_mma = 0;
    if (so_build_tc_image(s, img, &n_mma) == SO_OK) {
      rc = upload(&s->tc_image_dev, img.data(), img.size());
      s->tc_ok = (rc == SO_OK);
    }
  }
  if (rc == SO_OK) rc = upload(&s->Wst_dev, Wst.data(), (size_t)K * Hp);
  if (s->compact) {
    // the packed state serves the pipelined window kernels only: 17..20 hidden units, rectangular window
    if (!(Hp == 20 && s->win_ok)) {
      so_surrogate_destroy(s);
      so_set_error("the compact (24-bit) state needs 17..20 hidden units and a rectangular-window descriptor");
      return SO_EINVAL;
    }
    std::vector<int32_t> bias(Hp, 0);
    for (int j = 0; j < Hp; ++j) {
      long long lo = s->b1q_host[j];
      for (int k = 0; k < K; ++k) lo += std::min<int32_t>(s->W1q_host[(size_t)k * Hp + j], 0);
      bias[j] = (int32_t)(-lo);
    }
    if (rc == SO_OK) rc = upload(&s->bias24_dev, bias.data(), bias.size());
    std::vector<int32_t> W24((size_t)K * 4 * 8, 0);           // slot q = site p * 4 + part: weights of units 5*part .. 5*part+4
    for (int p = 0; p < K; ++p)
      for (int part = 0; part < 4; ++part)
        for (int i = 0; i < 5; ++i) W24[((size_t)p * 4 + part) * 8 + i] = Wst[(size_t)p * Hp + 5 * part + i];
    if (rc == SO_OK) rc = upload(&s->Wst24_dev, W24.data(), W24.size());
  }
  if (rc == SO_OK) rc = upload(&s->offs_dev, s->offs_host.data(), (size_t)K);
  if (rc == SO_OK) rc = upload(&s->W1q_dev, s->W1q_host.data(), (size_t)K * Hp);
  if (rc == SO_OK) rc = upload(&s->b1q_dev, s->b1q_host.data(), (size_t)Hp);
  if (rc == SO_OK) rc = upload(&s->w2q_dev, s->w2q_host.data(), (size_t)Hp);
  if (rc == SO_OK) rc = upload(&s->W1d_dev, W1, (size_t)K * H);
  if (rc == SO_OK) rc = upload(&s->b1d_dev, b1, (size_t)H);
  if (rc == SO_OK) rc = upload(&s->w2d_dev, W2, (size_t)H);
  if (rc == SO_OK && n_tree_nodes) {
    rc = upload(&s->t_feature_dev, feature, (size_t)n_tree_nodes);
    if (rc == SO_OK) rc = upload(&s->t_thr_dev, thr, (size_t)n_tree_nodes);
    if (rc == SO_OK) rc = upload(&s->t_left_dev, left, (size_t)n_tree_nodes);
    if (rc == SO_OK) rc = upload(&s->t_right_dev, right, (size_t)n_tree_nodes);
    if (rc == SO_OK) rc = upload(&s->t_active_dev, leaf_active, (size_t)n_tree_nodes);
    std::vector<int4> packed((size_t)n_tree_nodes);
    for (int i = 0; i < n_tree_nodes; ++i) {
      int4 nd;
      const bool leaf = feature[i] < 0;
      nd.x = leaf ? 0 : s->offs_host[feature[i]];
      nd.y = leaf ? 0 : left[i];
      nd.z = leaf ? 0 : right[i];
      nd.w = (leaf ? 1 : 0) | ((leaf && leaf_active[i]) ? 2 : 0) | ((!leaf && 0.0 <= thr[i]) ? 4 : 0) | ((!leaf && 1.0 <= thr[i]) ? 8 : 0) |
             (leaf ? 0 : (feature[i] << 8));      // bits 8..: the column tested
      packed[i] = nd;
    }
    if (rc == SO_OK) rc = upload(&s->t_packed_dev, packed.data(), (size_t)n_tree_nodes);
    // 8-byte nodes of the gated window kernel (so_sa_window_gated.cu): children by the value of the tested bit, the tested
    // column as patch-relative coordinates and as the window position its rows sit at
    const int Lw = s->win_b1 - s->win_a1 + 1, nseg = s->win_b2 - s->win_a2 + 1;
    if (rc == SO_OK && s->win_ok && n_tree_nodes <= 65535 && Lw <= 16 && nseg <= 16 && K <= 63) {
      std::vector<uint2> wt((size_t)n_tree_nodes);
      for (int i = 0; i < n_tree_nodes; ++i) {
        uint2 nd;
        if (feature[i] < 0) {
          nd.x = 0;
          nd.y = (1u | (leaf_active[i] ? 2u : 0u)) << 24;
        } else {
          const int o1 = offsets[2 * feature[i]], o2 = offsets[2 * feature[i] + 1];
          const unsigned c0 = (0.0 <= thr[i]) ? (unsigned)left[i] : (unsigned)right[i];     // bit 0: x = 0 <= thr ?
          const unsigned c1 = (1.0 <= thr[i]) ? (unsigned)left[i] : (unsigned)right[i];     // bit 1: x = 1 <= thr ?
          const unsigned pc = (unsigned)((s->win_b2 - o2) * Lw + (s->win_b1 - o1));
          nd.x = c0 | (c1 << 16);
          nd.y = (unsigned)(o1 - s->win_a1) | ((unsigned)(o2 - s->win_a2) << 8) | (pc << 16) | ((c0 != c1 ? 4u : 0u) << 24);
        }
        wt[i] = nd;
      }
      rc = upload((uint2 **)&s->wtree_dev, wt.data(), wt.size());
    }
  }
  if (rc != SO_OK) { so_surrogate_destroy(s); return rc; }
  *out = s;
  return SO_OK;
}

int so_surrogate_create(so_lattice *L, int K, int H, const int32_t *offsets, const double *W1, const double *b1,
                        const double *W2, double b2, double y_mean, double y_scale, int n_tree_nodes,
                        const int32_t *feature, const double *thr, const int32_t *left, const int32_t *right,
                        const uint8_t *leaf_active, so_surrogate **out) {
  return surrogate_create_impl(L, K, H, offsets, W1, b1, W2, b2, y_mean, y_scale, n_tree_nodes, feature, thr, left, right,
                               leaf_active, 30, out);
}

int so_surrogate_create_compact(so_lattice *L, int K, int H, const int32_t *offsets, const double *W1, const double *b1,
                                const double *W2, double b2, double y_mean, double y_scale, so_surrogate **out) {
  return surrogate_create_impl(L, K, H, offsets, W1, b1, W2, b2, y_mean, y_scale, 0, nullptr, nullptr, nullptr, nullptr,
                               nullptr, 22, out);
}

int so_surrogate_create_compact_gated(so_lattice *L, int K, int H, const int32_t *offsets, const double *W1, const double *b1,
                                      const double *W2, double b2, double y_mean, double y_scale, int n_tree_nodes,
                                      const int32_t *feature, const double *thr, const int32_t *left, const int32_t *right,
                                      const uint8_t *leaf_active, so_surrogate **out) {
  return surrogate_create_impl(L, K, H, offsets, W1, b1, W2, b2, y_mean, y_scale, n_tree_nodes, feature, thr, left, right,
                               leaf_active, 22, out);
}

int so_surrogate_destroy(so_surrogate *s) {
  if (!s) return SO_OK;
  cudaSetDevice(s->device);
  cudaFree(s->Wst24_dev); cudaFree(s->bias24_dev);
  cudaFree(s->tc_image_dev); cudaFree(s->Wst_dev); cudaFree(s->offs_dev); cudaFree(s->W1q_dev); cudaFree(s->b1q_dev); cudaFree(s->w2q_dev);
  cudaFree(s->W1d_dev); cudaFree(s->b1d_dev); cudaFree(s->w2d_dev);
  cudaFree(s->t_feature_dev); cudaFree(s->t_thr_dev); cudaFree(s->t_left_dev); cudaFree(s->t_right_dev);
  cudaFree(s->t_active_dev); cudaFree(s->t_packed_dev); cudaFree(s->wtree_dev);
  delete s;
  return SO_OK;
}

int so_surrogate_info(const so_surrogate *s, int *K, int *H, int *Hp, int *e_h, int *e_w, int *gated) {
  SO_REQUIRE(s, "surrogate is NULL");
  if (K) *K = s->K;
  if (H) *H = s->H;
  if (Hp) *Hp = s->Hp;
  if (e_h) *e_h = s->e_h;
  if (e_w) *e_w = s->e_w;
  if (gated) *gated = s->gated;
  return SO_OK;
}

// CUtensorMap over `base` viewed as int32 [n_chains][d][d * site_words], box = one window (Lw sites x nseg rows); 1 = ok
static int window_tensor_map(CUtensorMap *map, void *base, const so_lattice *L, int64_t n_chains, int site_words, int Lw,
                             int nseg) {
  if ((uint64_t)Lw * site_words > 256 || nseg > 256) return 0;
  typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                               const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  int ok = 0;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn &&
      qres == cudaDriverEntryPointSuccess) {
    cuuint64_t gdim[3] = {(cuuint64_t)L->d * site_words, (cuuint64_t)L->d, (cuuint64_t)n_chains};
    cuuint64_t gstr[2] = {(cuuint64_t)L->d * site_words * 4, (cuuint64_t)L->N * site_words * 4};
    cuuint32_t box[3] = {(cuuint32_t)(Lw * site_words), (cuuint32_t)nseg, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult cr = ((EncodeFn)fn)(map, CU_TENSOR_MAP_DATA_TYPE_INT32, 3, base, gdim, gstr, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ok = (cr == CUDA_SUCCESS) ? 1 : 0;
  }
  (void)cudaGetLastError();
  return ok;
}

int so_chains_attach(so_chains *c, so_surrogate *s) {
  SO_REQUIRE(c && s, "NULL argument");
  SO_REQUIRE(s->lat == c->lat, "surrogate and chains were built on different lattices");
  const so_lattice *L = c->lat;
  SO_CUDA(cudaSetDevice(L->device));
  cudaFree(c->hq_dev); cudaFree(c->hi_dev); cudaFree(c->lo_dev); cudaFree(c->nact_dev);
  c->hq_dev = nullptr; c->hi_dev = c->lo_dev = nullptr; c->nact_dev = nullptr; c->sur = nullptr;
  const int site_words = s->compact ? 16 : s->Hp;          // uint32 per site in memory
  size_t bytes = (size_t)c->n_chains * L->N * site_words * 4;
  cudaError_t e = cudaMalloc(&c->hq_dev, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&c->hi_dev, (size_t)c->n_chains * 8);
  if (e == cudaSuccess) e = cudaMalloc(&c->lo_dev, (size_t)c->n_chains * 8);
  if (e == cudaSuccess) e = cudaMalloc(&c->nact_dev, (size_t)c->n_chains * 4);
  if (e != cudaSuccess) {
    so_set_error("surrogate state allocation (%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    cudaFree(c->hq_dev); cudaFree(c->hi_dev); cudaFree(c->lo_dev); cudaFree(c->nact_dev);
    c->hq_dev = nullptr; c->hi_dev = c->lo_dev = nullptr; c->nact_dev = nullptr;
    (void)cudaGetLastError();
    return SO_ENOMEM;
  }
  c->sur = s;
  // tensor map over hq viewed as int32 [n_chains][d][d*site_words]; one box = the window of a flip (so_sa_window.cu).  Gated
  // window descriptors also get the per-row records of so_sa_window_gated.cu (16 bytes per row) and their map.
  c->tmap_ok = 0;
  c->tmap_meta_ok = 0;
  c->meta_valid = 0;
  cudaFree(c->gate_meta_dev);
  c->gate_meta_dev = nullptr;
  if (s->win_ok) {
    const int Lw = s->win_b1 - s->win_a1 + 1, nseg = s->win_b2 - s->win_a2 + 1;
    c->tmap_ok = window_tensor_map(&c->tmap, c->hq_dev, L, c->n_chains, site_words, Lw, nseg);
    if (s->gated && s->wtree_dev && s->HP4 == 5 && c->tmap_ok) {          // (the shapes so_window_gated_ok accepts)
      if (cudaMalloc(&c->gate_meta_dev, (size_t)c->n_chains * L->N * 16) == cudaSuccess) {
        c->tmap_meta_ok = window_tensor_map(&c->tmap_meta, c->gate_meta_dev, L, c->n_chains, 4, Lw, nseg);
      } else {
        c->gate_meta_dev = nullptr;
        (void)cudaGetLastError();                  // no records: the cached-gate kernel with its own scratch serves the gate
      }
    }
  }
  if (s->compact && s->gated && !so_window_gated_ok(c)) {
    c->sur = nullptr;
    so_set_error("the compact (24-bit) gated state is served by the gated window kernel only, which cannot run this shape");
    return SO_EINVAL;
  }
  SO_TRY(rescore(c, 0, c->n_chains));
  SO_CUDA(cudaStreamSynchronize(c->stream));
  return SO_OK;
}

int so_objective(so_chains *c, int64_t chain0, int64_t n, double *of_out, int64_t *hi, int64_t *lo, int32_t *n_active) {
  SO_TRY(check_range(c, chain0, n));
  SO_REQUIRE(c->sur, "no surrogate attached");
  if (n == 0) return SO_OK;
  SO_CUDA(cudaSetDevice(c->lat->device));
  if (of_out) {
    DevBuf of;
    SO_TRY(of.alloc((size_t)n * 8));
    SO_TRY(so_launch_objective(c->sur, c->hi_dev + chain0, c->lo_dev + chain0, c->nact_dev + chain0, n, of.as<double>(),
                               c->stream));
    SO_CUDA(cudaMemcpyAsync(of_out, of.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    SO_CUDA(cudaStreamSynchronize(c->stream));
  }
  if (hi) SO_CUDA(cudaMemcpy(hi, c->hi_dev + chain0, (size_t)n * 8, cudaMemcpyDeviceToHost));
  if (lo) SO_CUDA(cudaMemcpy(lo, c->lo_dev + chain0, (size_t)n * 8, cudaMemcpyDeviceToHost));
  if (n_active) SO_CUDA(cudaMemcpy(n_active, c->nact_dev + chain0, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return SO_OK;
}

int so_score(so_chains *c, int64_t chain0, int64_t n, double *of_out) {
  SO_TRY(check_range(c, chain0, n));
  SO_REQUIRE(c->sur, "no surrogate attached");
  SO_CUDA(cudaSetDevice(c->lat->device));
  SO_TRY(rescore(c, chain0, n));
  SO_CUDA(cudaStreamSynchronize(c->stream));
  if (of_out) return so_objective(c, chain0, n, of_out, nullptr, nullptr, nullptr);
  return SO_OK;
}

int so_get_preact(so_chains *c, int64_t chain, int32_t *hq) {
  SO_TRY(check_range(c, chain, 1));
  SO_REQUIRE(c->sur && hq, "no surrogate attached or NULL output");
  SO_CUDA(cudaSetDevice(c->lat->device));
  size_t row = (size_t)c->lat->N * c->sur->Hp;
  if (c->sur->compact) {
    DevBuf tmp;
    SO_TRY(tmp.alloc(row * 4));
    SO_TRY(so_launch_unpack24((const char *)c->hq_dev + (size_t)chain * c->lat->N * 64, tmp.as<int32_t>(), c->lat->N,
                              c->sur->bias24_dev, c->stream));
    SO_CUDA(cudaMemcpyAsync(hq, tmp.p, row * 4, cudaMemcpyDeviceToHost, c->stream));
    SO_CUDA(cudaStreamSynchronize(c->stream));
    return SO_OK;
  }
  SO_CUDA(cudaMemcpy(hq, c->hq_dev + (size_t)chain * row, row * 4, cudaMemcpyDeviceToHost));
  return SO_OK;
}

int so_score_batch(so_lattice *L, so_surrogate *s, int64_t n, const uint32_t *bits, double *rate, int32_t *lsc_hist,
                   int32_t *nh3_exp) {
  SO_REQUIRE(L && bits, "NULL argument");
  SO_REQUIRE(n >= 0, "negative batch");
  SO_REQUIRE(!rate || s, "rate requested without a surrogate");
  SO_REQUIRE(!s || s->lat == L, "surrogate built on a different lattice");
  if (n == 0) return SO_OK;
  SO_CUDA(cudaSetDevice(L->device));
  const int64_t slab = std::min<int64_t>(n, 1 << 20);
  DevBuf b_bits, b_hi, b_lo, b_n, b_of, b_lh, b_ne;
  SO_TRY(b_bits.alloc((size_t)slab * L->W * 4));
  if (rate) {
    SO_TRY(b_hi.alloc((size_t)slab * 8)); SO_TRY(b_lo.alloc((size_t)slab * 8));
    SO_TRY(b_n.alloc((size_t)slab * 4)); SO_TRY(b_of.alloc((size_t)slab * 8));
  }
  if (lsc_hist) SO_TRY(b_lh.alloc((size_t)slab * 12));
  if (nh3_exp) SO_TRY(b_ne.alloc((size_t)slab * 32));
  cudaStream_t st = 0;
  for (int64_t c0 = 0; c0 < n; c0 += slab) {
    int64_t nc = std::min(slab, n - c0);
    SO_CUDA(cudaMemcpyAsync(b_bits.p, bits + c0 * L->W, (size_t)nc * L->W * 4, cudaMemcpyHostToDevice, st));
    if (rate) {
      SO_TRY(so_launch_score(L, s, b_bits.as<uint32_t>(), nc, nullptr, b_hi.as<long long>(), b_lo.as<long long>(),
                             b_n.as<int32_t>(), st));
      SO_TRY(so_launch_objective(s, b_hi.as<long long>(), b_lo.as<long long>(), b_n.as<int32_t>(), nc, b_of.as<double>(),
                                 st));
      SO_CUDA(cudaMemcpyAsync(rate + c0, b_of.p, (size_t)nc * 8, cudaMemcpyDeviceToHost, st));
    }
    if (lsc_hist || nh3_exp) {
      SO_TRY(so_launch_descriptors(L, b_bits.as<uint32_t>(), nc, nullptr, nullptr, nullptr, lsc_hist ? b_lh.as<int32_t>() : nullptr,
                                   nullptr, nh3_exp ? b_ne.as<int32_t>() : nullptr, st));
      if (lsc_hist) SO_CUDA(cudaMemcpyAsync(lsc_hist + c0 * 3, b_lh.p, (size_t)nc * 12, cudaMemcpyDeviceToHost, st));
      if (nh3_exp) SO_CUDA(cudaMemcpyAsync(nh3_exp + c0 * 8, b_ne.p, (size_t)nc * 32, cudaMemcpyDeviceToHost, st));
    }
    SO_CUDA(cudaStreamSynchronize(st));
  }
  return SO_OK;
}

int so_eval_rows(so_surrogate *s, int64_t n_rows, const uint8_t *X, uint8_t *active, double *rate) {
  SO_REQUIRE(s && (X || n_rows == 0), "NULL argument");
  SO_REQUIRE(n_rows >= 0, "negative row count");
  if (n_rows == 0) return SO_OK;
  SO_CUDA(cudaSetDevice(s->lat->device));
  DevBuf bx, ba, br;
  SO_TRY(bx.alloc((size_t)n_rows * s->K));
  if (active) SO_TRY(ba.alloc((size_t)n_rows));
  if (rate) SO_TRY(br.alloc((size_t)n_rows * 8));
  SO_CUDA(cudaMemcpy(bx.p, X, (size_t)n_rows * s->K, cudaMemcpyHostToDevice));
  SO_TRY(so_launch_eval_rows(s, n_rows, bx.as<uint8_t>(), ba.as<uint8_t>(), br.as<double>(), 0));
  if (active) SO_CUDA(cudaMemcpy(active, ba.p, (size_t)n_rows, cudaMemcpyDeviceToHost));
  if (rate) SO_CUDA(cudaMemcpy(rate, br.p, (size_t)n_rows * 8, cudaMemcpyDeviceToHost));
  SO_CUDA(cudaDeviceSynchronize());
  return SO_OK;
}

// ---- annealing --------------------------------------------------------------------------------------------
int so_anneal(so_chains *c, const so_anneal_args *a, so_anneal_stats *stats) {
  SO_REQUIRE(c && a, "NULL argument");
  SO_REQUIRE(c->sur, "no surrogate attached");   // SO_ESTATE would also do; keep the message
  SO_REQUIRE(a->n_steps >= 0 && a->step0 >= 0, "negative step range");
  SO_REQUIRE(a->temps || a->n_steps == 0, "temps is NULL");
  SO_REQUIRE((a->proposals == nullptr) == (a->uniforms == nullptr), "proposals and uniforms must be given together");
  NvtxRange nvtx_range("so_anneal (sweep)");
  const so_lattice *L = c->lat;
  if (stats) memset(stats, 0, sizeof(*stats));
  if (a->n_steps == 0) return SO_OK;
  SO_CUDA(cudaSetDevice(L->device));
  const size_t ns = (size_t)a->n_steps, nc = (size_t)c->n_chains;
  if (a->proposals)
    for (size_t i = 0; i < nc * ns; ++i)
      SO_REQUIRE(a->proposals[i] >= 0 && a->proposals[i] < L->N, "proposal %zu = %d out of range", i, a->proposals[i]);
  DevBuf d_prop, d_uni, d_acc;
  if (c->temps_cap < ns * 8) {
    cudaFree(c->temps_dev);
    c->temps_dev = nullptr;
    c->temps_cap = 0;
    const size_t want = std::max<size_t>(ns * 8, 1 << 16);
    if (cudaMalloc(&c->temps_dev, want) != cudaSuccess) {
      cudaGetLastError();
      so_set_error("cudaMalloc(%zu bytes) failed for the temperature schedule", want);
      return SO_ENOMEM;
    }
    c->temps_cap = want;
  }
  SO_CUDA(cudaMemcpyAsync(c->temps_dev, a->temps, ns * 8, cudaMemcpyHostToDevice, c->stream));
  if (a->proposals) {
    SO_TRY(d_prop.alloc(nc * ns * 4));
    SO_TRY(d_uni.alloc(nc * ns * 4));
    SO_CUDA(cudaMemcpyAsync(d_prop.p, a->proposals, nc * ns * 4, cudaMemcpyHostToDevice, c->stream));
    SO_CUDA(cudaMemcpyAsync(d_uni.p, a->uniforms, nc * ns * 4, cudaMemcpyHostToDevice, c->stream));
  }
  if (a->accept_out) SO_TRY(d_acc.alloc(nc * ns));
  SO_CUDA(cudaMemsetAsync(c->counters_dev, 0, 3 * sizeof(unsigned long long), c->stream));
  SaParams P;
  memset(&P, 0, sizeof(P));
  P.type_hist = c->type_hist_dev;
  {
    // resident kernel, difference pass: a lane per whole row pays off for the full translation descriptor without a
    // gate (+12 % at d = 12, K = 144); gated (scattered active rows) and 7x7 windows keep 32/HP4 rows per pass
    // (profiles/r02_resident_rowlane_ab.txt).  SO_RES_ROWLANE = 0 / 1 forces either.
    const char *e = getenv("SO_RES_ROWLANE");
    P.variant_flags = e ? (atoi(e) & 1) : ((!c->sur->gated && c->sur->K >= 96) ? 1 : 0);
    // bit 1: gated FULL translation descriptor (column k <-> offset (k % d, k / d)): persistent list of active rows
    const so_surrogate *su = c->sur;
    bool full = su->gated && su->K == L->N && (int)su->offs_host.size() == su->K && L->d <= 255;
    for (int k = 0; full && k < su->K; ++k)
      full = su->offs_host[k] == (int32_t)(((k / L->d) << 16) | (k % L->d));
    const char *e2 = getenv("SO_RES_ACTIVE_LIST");
    if (full && !(e2 && e2[0] == '0')) {
      P.variant_flags |= 2;
      // ... then a lane per whole row wins the difference pass too (golden surrogate +5 %, 60 % active rows +17 %)
      if (!e) P.variant_flags |= 1;
      // (tried and dropped: walks on a doubled bit plane + look-up tables instead of wrap() arithmetic remove ~250 of the
      //  1,310 instructions of a gated step and change nothing -- 5.26e8 vs 5.39e8: with one warp per chain the step is a
      //  serial chain of phases, 5.7 cycles per instruction at 16 warps per SM, not an issue-bound loop)
    }
  }
  P.d = L->d; P.N = L->N; P.W = L->W; P.n_chains = c->n_chains;
  P.bits = c->bits_dev; P.hq = c->hq_dev; P.hi = c->hi_dev; P.lo = c->lo_dev; P.nact = c->nact_dev;
  P.tscale = c->tscale_dev; P.temps = c->temps_dev;
  P.step0 = a->step0; P.n_steps = a->n_steps; P.seed = a->seed; P.chain_id0 = a->chain_id0;
  P.proposals = a->proposals ? d_prop.as<int32_t>() : nullptr;
  P.uniforms = a->uniforms ? d_uni.as<float>() : nullptr;
  P.accept_out = a->accept_out ? d_acc.as<uint8_t>() : nullptr;
  P.exact_guard = a->exact_guard; P.guard_tol = a->guard_tol; P.variant = a->variant;
  P.counters = c->counters_dev;
  P.gate_scratch = nullptr;
  P.sv = make_view(c->sur);
  SO_CUDA(cudaEventRecord(c->ev0, c->stream));
  int kernel_id = SO_K_NONE;
  SO_TRY(so_launch_anneal(c, P, c->stream, &kernel_id));
  SO_CUDA(cudaEventRecord(c->ev1, c->stream));
  if (a->accept_out) SO_CUDA(cudaMemcpyAsync(a->accept_out, d_acc.p, nc * ns, cudaMemcpyDeviceToHost, c->stream));
  SO_CUDA(cudaStreamSynchronize(c->stream));
  if (stats) {
    unsigned long long cnt[3];
    SO_CUDA(cudaMemcpy(cnt, c->counters_dev, sizeof(cnt), cudaMemcpyDeviceToHost));
    stats->flips = (int64_t)(nc * ns);
    stats->accepted = (int64_t)cnt[1];
    stats->guard_fired = (int64_t)cnt[2];
    SO_CUDA(cudaEventElapsedTime(&stats->kernel_ms, c->ev0, c->ev1));
    stats->kernel_id = kernel_id;
  }
  return SO_OK;
}

const char *so_kernel_name(int kernel_id) {
  switch (kernel_id) {
    case SO_K_TEAM: return "sa_team_kernel";
    case SO_K_RESIDENT: return "sa_resident_kernel";
    case SO_K_GATE_INDEX: return "sa_resident_kernel<RES=false> (cached gates, state in HBM)";
    case SO_K_WINDOW: return "sa_window_kernel";
    case SO_K_WINDOW_GATED: return "sa_window_gated_kernel";
    case SO_K_FLAT: return "sa_flat_kernel";
    case SO_K_GATED: return "sa_gated_kernel";
    case SO_K_ANALYTICAL: return "lsc_anneal_kernel";
    default: return "none";
  }
}

int so_set_tscale(so_chains *c, int64_t chain0, int64_t n, const double *tscale) {
  SO_TRY(check_range(c, chain0, n));
  SO_REQUIRE(tscale || n == 0, "tscale is NULL");
  SO_CUDA(cudaSetDevice(c->lat->device));
  SO_CUDA(cudaMemcpy(c->tscale_dev + chain0, tscale, (size_t)n * 8, cudaMemcpyHostToDevice));
  return SO_OK;
}

int so_get_tscale(so_chains *c, int64_t chain0, int64_t n, double *tscale) {
  SO_TRY(check_range(c, chain0, n));
  SO_REQUIRE(tscale || n == 0, "tscale is NULL");
  SO_CUDA(cudaSetDevice(c->lat->device));
  SO_CUDA(cudaMemcpy(tscale, c->tscale_dev + chain0, (size_t)n * 8, cudaMemcpyDeviceToHost));
  return SO_OK;
}

// ---- type histograms carried through the annealing (SURVEY 9.5 frozen step order: commit x, h, OF, descriptors) --------
int so_track_types(so_chains *c, int enable) {
  SO_REQUIRE(c, "chains handle is NULL");
  SO_CUDA(cudaSetDevice(c->lat->device));
  if (!enable) {
    if (c->type_hist_dev) cudaFree(c->type_hist_dev);
    c->type_hist_dev = nullptr;
    return SO_OK;
  }
  if (!c->type_hist_dev) {
    const size_t bytes = (size_t)c->n_chains * 11 * 4;
    if (cudaMalloc((void **)&c->type_hist_dev, std::max<size_t>(bytes, 16)) != cudaSuccess) {
      cudaGetLastError();
      c->type_hist_dev = nullptr;
      so_set_error("cudaMalloc(%zu bytes) failed for the tracked type histograms", bytes);
      return SO_ENOMEM;
    }
  }
  SO_TRY(refresh_types(c, 0, c->n_chains));
  SO_CUDA(cudaStreamSynchronize(c->stream));
  return SO_OK;
}

int so_type_histograms(so_chains *c, int64_t chain0, int64_t n, int32_t *lsc_hist, int32_t *nh3_exp) {
  SO_TRY(check_range(c, chain0, n));
  SO_REQUIRE(c->type_hist_dev, "type tracking is off (so_track_types)");
  if (n == 0) return SO_OK;
  SO_CUDA(cudaSetDevice(c->lat->device));
  if (nh3_exp)
    SO_CUDA(cudaMemcpyAsync(nh3_exp, c->type_hist_dev + chain0 * 8, (size_t)n * 32, cudaMemcpyDeviceToHost, c->stream));
  if (lsc_hist)
    SO_CUDA(cudaMemcpyAsync(lsc_hist, c->type_hist_dev + c->n_chains * 8 + chain0 * 3, (size_t)n * 12, cudaMemcpyDeviceToHost,
                            c->stream));
  SO_CUDA(cudaStreamSynchronize(c->stream));
  return SO_OK;
}

// ---- checkpoint / resume ------------------------------------------------------------------------------------
// A chain's whole state is a function of its bits (the first-layer state and the sums are exact integers, recomputed
// by the state build) plus the temperature scale; with counter-based draws a run resumed from (bits, tscale, step,
// seed) continues bit for bit.  File: 64-byte header, bits[n][W] u32, tscale[n] f64, hi[n] i64, lo[n] i64, nact[n] i32
// (the sums are stored only to verify the rebuilt state on load).
namespace {
struct CkptHeader {
  char magic[8];            // "SOB200CK"
  uint32_t version, d, N, W;
  int64_t n_chains;
  int64_t step;
  uint64_t seed;
  int32_t has_sums, K, Hp, pad;
};
static_assert(sizeof(CkptHeader) == 64, "checkpoint header layout");
}  // namespace

int so_chains_save(so_chains *c, const char *path, int64_t step, uint64_t seed) {
  SO_REQUIRE(c && path, "NULL argument");
  const so_lattice *L = c->lat;
  SO_CUDA(cudaSetDevice(L->device));
  SO_CUDA(cudaStreamSynchronize(c->stream));
  const size_t n = (size_t)c->n_chains;
  std::vector<uint32_t> bits(n * L->W);
  std::vector<double> ts(n);
  std::vector<long long> hi(n), lo(n);
  std::vector<int32_t> na(n);
  SO_CUDA(cudaMemcpy(bits.data(), c->bits_dev, bits.size() * 4, cudaMemcpyDeviceToHost));
  SO_CUDA(cudaMemcpy(ts.data(), c->tscale_dev, n * 8, cudaMemcpyDeviceToHost));
  CkptHeader h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, "SOB200CK", 8);
  h.version = 1; h.d = (uint32_t)L->d; h.N = (uint32_t)L->N; h.W = (uint32_t)L->W;
  h.n_chains = c->n_chains; h.step = step; h.seed = seed;
  if (c->sur) {
    h.has_sums = 1; h.K = c->sur->K; h.Hp = c->sur->Hp;
    SO_CUDA(cudaMemcpy(hi.data(), c->hi_dev, n * 8, cudaMemcpyDeviceToHost));
    SO_CUDA(cudaMemcpy(lo.data(), c->lo_dev, n * 8, cudaMemcpyDeviceToHost));
    SO_CUDA(cudaMemcpy(na.data(), c->nact_dev, n * 4, cudaMemcpyDeviceToHost));
  }
  FILE *f = fopen(path, "wb");
  SO_REQUIRE(f, "cannot open %s for writing", path);
  bool ok = fwrite(&h, sizeof(h), 1, f) == 1 && fwrite(bits.data(), 4, bits.size(), f) == bits.size() &&
            fwrite(ts.data(), 8, n, f) == n;
  if (ok && h.has_sums)
    ok = fwrite(hi.data(), 8, n, f) == n && fwrite(lo.data(), 8, n, f) == n && fwrite(na.data(), 4, n, f) == n;
  ok = (fclose(f) == 0) && ok;
  SO_REQUIRE(ok, "short write to %s", path);
  return SO_OK;
}

int so_chains_load(so_chains *c, const char *path, int64_t *step, uint64_t *seed) {
  SO_REQUIRE(c && path, "NULL argument");
  const so_lattice *L = c->lat;
  FILE *f = fopen(path, "rb");
  SO_REQUIRE(f, "cannot open %s", path);
  CkptHeader h;
  bool ok = fread(&h, sizeof(h), 1, f) == 1;
  if (!ok || memcmp(h.magic, "SOB200CK", 8) != 0 || h.version != 1) {
    fclose(f);
    so_set_error("%s is not a chain checkpoint of this library", path);
    return SO_EINVAL;
  }
  if ((int)h.d != L->d || h.n_chains != c->n_chains || (int)h.W != L->W) {
    fclose(f);
    so_set_error("checkpoint holds %lld chains on a %ux%u lattice; this handle has %lld chains on %dx%d",
                 (long long)h.n_chains, h.d, h.d, (long long)c->n_chains, L->d, L->d);
    return SO_EINVAL;
  }
  const size_t n = (size_t)c->n_chains;
  std::vector<uint32_t> bits(n * L->W);
  std::vector<double> ts(n);
  std::vector<long long> hi(n), lo(n);
  std::vector<int32_t> na(n);
  ok = fread(bits.data(), 4, bits.size(), f) == bits.size() && fread(ts.data(), 8, n, f) == n;
  if (ok && h.has_sums)
    ok = fread(hi.data(), 8, n, f) == n && fread(lo.data(), 8, n, f) == n && fread(na.data(), 4, n, f) == n;
  fclose(f);
  SO_REQUIRE(ok, "%s is truncated", path);
  SO_TRY(so_set_occupancy_bits(c, 0, c->n_chains, bits.data()));          // validates the bits, rebuilds the state
  SO_TRY(so_set_tscale(c, 0, c->n_chains, ts.data()));
  if (h.has_sums && c->sur && h.K == c->sur->K && h.Hp == c->sur->Hp) {   // the rebuilt sums must be the saved ones
    std::vector<long long> hi2(n), lo2(n);
    std::vector<int32_t> na2(n);
    SO_CUDA(cudaMemcpy(hi2.data(), c->hi_dev, n * 8, cudaMemcpyDeviceToHost));
    SO_CUDA(cudaMemcpy(lo2.data(), c->lo_dev, n * 8, cudaMemcpyDeviceToHost));
    SO_CUDA(cudaMemcpy(na2.data(), c->nact_dev, n * 4, cudaMemcpyDeviceToHost));
    SO_REQUIRE(hi2 == hi && lo2 == lo && na2 == na,
               "checkpoint %s was written with a different surrogate (objective sums do not reproduce)", path);
  }
  if (step) *step = h.step;
  if (seed) *seed = h.seed;
  return SO_OK;
}

// ---- LSC analytical objective ------------------------------------------------------------------------------
int so_lsc_analytical(so_chains *c, int64_t chain0, int64_t n, double *site_rates, double *struct_rate) {
  SO_TRY(check_range(c, chain0, n));
  const so_lattice *L = c->lat;
  SO_REQUIRE(L->N <= so_lsc_max_sites(), "the analytical objective is limited to %d sites (this lattice has %d)",
             so_lsc_max_sites(), L->N);
  if (n == 0) return SO_OK;
  SO_CUDA(cudaSetDevice(L->device));
  DevBuf d_sr, d_rate;
  if (site_rates) SO_TRY(d_sr.alloc((size_t)n * L->N * 8));
  if (struct_rate) SO_TRY(d_rate.alloc((size_t)n * 8));
  SO_TRY(so_launch_lsc_analytical(L, c->bits_dev + chain0 * L->W, n, d_sr.as<double>(), d_rate.as<double>(), c->stream));
  if (site_rates) SO_CUDA(cudaMemcpyAsync(site_rates, d_sr.p, (size_t)n * L->N * 8, cudaMemcpyDeviceToHost, c->stream));
  if (struct_rate) SO_CUDA(cudaMemcpyAsync(struct_rate, d_rate.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
  SO_CUDA(cudaStreamSynchronize(c->stream));
  return SO_OK;
}

int so_anneal_analytical(so_chains *c, const so_anneal_args *a, so_anneal_stats *stats, double *of_out) {
  SO_REQUIRE(c && a, "NULL argument");
  const so_lattice *L = c->lat;
  SO_REQUIRE(L->N <= so_lsc_max_sites(), "the analytical objective is limited to %d sites (this lattice has %d)",
             so_lsc_max_sites(), L->N);
  SO_REQUIRE(a->n_steps >= 0 && a->step0 >= 0, "negative step range");
  SO_REQUIRE(a->temps || a->n_steps == 0, "temps is NULL");
  SO_REQUIRE((a->proposals == nullptr) == (a->uniforms == nullptr), "proposals and uniforms must be given together");
  if (stats) memset(stats, 0, sizeof(*stats));
  SO_CUDA(cudaSetDevice(L->device));
  const size_t ns = (size_t)a->n_steps, nc = (size_t)c->n_chains;
  if (a->proposals)
    for (size_t i = 0; i < nc * ns; ++i)
      SO_REQUIRE(a->proposals[i] >= 0 && a->proposals[i] < L->N, "proposal %zu = %d out of range", i, a->proposals[i]);
  DevBuf d_temps, d_prop, d_uni, d_acc, d_of;
  SO_TRY(d_temps.alloc(ns * 8));
  if (ns) SO_CUDA(cudaMemcpyAsync(d_temps.p, a->temps, ns * 8, cudaMemcpyHostToDevice, c->stream));
  if (a->proposals) {
    SO_TRY(d_prop.alloc(nc * ns * 4));
    SO_TRY(d_uni.alloc(nc * ns * 4));
    SO_CUDA(cudaMemcpyAsync(d_prop.p, a->proposals, nc * ns * 4, cudaMemcpyHostToDevice, c->stream));
    SO_CUDA(cudaMemcpyAsync(d_uni.p, a->uniforms, nc * ns * 4, cudaMemcpyHostToDevice, c->stream));
  }
  if (a->accept_out) SO_TRY(d_acc.alloc(nc * ns));
  SO_TRY(d_of.alloc(nc * 8));
  SO_CUDA(cudaMemsetAsync(c->counters_dev, 0, 3 * sizeof(unsigned long long), c->stream));
  SaParams P;
  memset(&P, 0, sizeof(P));
  P.d = L->d; P.N = L->N; P.W = L->W; P.n_chains = c->n_chains;
  P.bits = c->bits_dev; P.tscale = c->tscale_dev; P.temps = d_temps.as<double>();
  P.step0 = a->step0; P.n_steps = a->n_steps; P.seed = a->seed; P.chain_id0 = a->chain_id0;
  P.proposals = a->proposals ? d_prop.as<int32_t>() : nullptr;
  P.uniforms = a->uniforms ? d_uni.as<float>() : nullptr;
  P.accept_out = a->accept_out ? d_acc.as<uint8_t>() : nullptr;
  P.counters = c->counters_dev;
  SO_CUDA(cudaEventRecord(c->ev0, c->stream));
  SO_TRY(so_launch_anneal_analytical(c, P, d_of.as<double>(), c->stream));
  SO_CUDA(cudaEventRecord(c->ev1, c->stream));
  if (a->accept_out && ns) SO_CUDA(cudaMemcpyAsync(a->accept_out, d_acc.p, nc * ns, cudaMemcpyDeviceToHost, c->stream));
  if (of_out) SO_CUDA(cudaMemcpyAsync(of_out, d_of.p, nc * 8, cudaMemcpyDeviceToHost, c->stream));
  SO_TRY(rescore(c, 0, c->n_chains));               // keep an attached surrogate state consistent with the new bits
  SO_CUDA(cudaStreamSynchronize(c->stream));
  if (stats) {
    unsigned long long cnt[3];
    SO_CUDA(cudaMemcpy(cnt, c->counters_dev, sizeof(cnt), cudaMemcpyDeviceToHost));
    stats->flips = (int64_t)(nc * ns);
    stats->accepted = (int64_t)cnt[1];
    SO_CUDA(cudaEventElapsedTime(&stats->kernel_ms, c->ev0, c->ev1));
    stats->kernel_id = SO_K_ANALYTICAL;
  }
  return SO_OK;
}

// ---- best structure / replica exchange -----------------------------------------------------------------------
int so_best_record_dev(so_chains *c, int64_t chain_id0, void *record_dev, void *stream) {
  SO_REQUIRE(c && record_dev, "NULL argument");
  SO_REQUIRE(c->sur, "no surrogate attached");
  NvtxRange nvtx_range("best-structure record (all-gather send buffer)");
  SO_CUDA(cudaSetDevice(c->lat->device));
  SO_TRY(so_launch_best(c, chain_id0, record_dev, (cudaStream_t)stream));
  return SO_OK;
}

int so_best(so_chains *c, double *of_out, int64_t *chain_out, uint32_t *bits) {
  SO_REQUIRE(c, "chains handle is NULL");
  SO_REQUIRE(c->sur, "no surrogate attached");
  const so_lattice *L = c->lat;
  SO_CUDA(cudaSetDevice(L->device));
  DevBuf rec;
  size_t bytes = 16 + (size_t)L->W * 4;
  SO_TRY(rec.alloc(bytes));
  SO_TRY(so_launch_best(c, 0, rec.p, c->stream));
  std::vector<unsigned char> host(bytes);
  SO_CUDA(cudaMemcpyAsync(host.data(), rec.p, bytes, cudaMemcpyDeviceToHost, c->stream));
  SO_CUDA(cudaStreamSynchronize(c->stream));
  if (of_out) memcpy(of_out, host.data(), 8);
  if (chain_out) memcpy(chain_out, host.data() + 8, 8);
  if (bits) memcpy(bits, host.data() + 16, (size_t)L->W * 4);
  return SO_OK;
}

int so_objectives_dev(so_chains *c, double *of_dev, void *stream) {
  SO_REQUIRE(c && of_dev, "NULL argument");
  SO_REQUIRE(c->sur, "no surrogate attached");
  SO_CUDA(cudaSetDevice(c->lat->device));
  return so_launch_objective(c->sur, c->hi_dev, c->lo_dev, c->nact_dev, c->n_chains, of_dev, (cudaStream_t)stream);
}

int so_exchange(so_chains *c, int rank, int world, int ladder, int parity, uint64_t seed, int64_t round, double t_ref,
                const double *of_all_dev, double *tscale_all_dev, void *stream, int64_t *n_swapped) {
  SO_REQUIRE(c && of_all_dev && tscale_all_dev, "NULL argument");
  SO_REQUIRE(ladder >= 2 && ladder <= 32, "ladder size %d outside [2, 32]", ladder);
  SO_REQUIRE(parity == 0 || parity == 1, "parity must be 0 or 1");
  SO_REQUIRE(world >= 1 && rank >= 0 && rank < world, "rank %d outside world %d", rank, world);
  NvtxRange nvtx_range("replica-exchange decisions");
  const int64_t n_global = c->n_chains * (int64_t)world;
  SO_REQUIRE(n_global % ladder == 0, "world*n_chains=%lld is not a multiple of the ladder size %d", (long long)n_global,
             ladder);
  SO_REQUIRE(t_ref > 0.0, "t_ref must be positive");
  SO_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (c->xch_cap < n_global) {
    cudaFree(c->xch_scratch);
    c->xch_scratch = nullptr; c->xch_cap = 0;
    cudaError_t e = cudaMalloc(&c->xch_scratch, (size_t)n_global * 8);
    if (e != cudaSuccess) { so_set_error("cudaMalloc failed: %s", cudaGetErrorString(e)); return SO_ENOMEM; }
    c->xch_cap = n_global;
  }
  SO_CUDA(cudaMemsetAsync(c->counters_dev + 4, 0, sizeof(unsigned long long), st));
  SO_TRY(so_launch_exchange(c, rank, world, ladder, parity, seed, round, t_ref, of_all_dev, tscale_all_dev,
                            c->xch_scratch, c->counters_dev + 4, st));
  unsigned long long ns = 0;
  SO_CUDA(cudaMemcpyAsync(&ns, c->counters_dev + 4, sizeof(ns), cudaMemcpyDeviceToHost, st));
  SO_CUDA(cudaStreamSynchronize(st));
  if (n_swapped) *n_swapped = (int64_t)ns;
  return SO_OK;
}

}  // extern "C"
