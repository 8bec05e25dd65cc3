// so_lsc.cu -- the analytical objective of the LSC toy model on the device.
//
// Replaces  LSC_cat.eval_struc_rate_anal / compute_site_rates_lsc   OML/LSC_cat.py:248-316
//           optimize(mode='analytical') hot loop                    OML/optimize_SA.py:59-132 (objective at :104)
// One thread block per structure: site typing from the bit-plane (closed-form rules of so_desc.cu), the typed sites
// enumerated in ascending order (KMC_lat.site_type_inds), the dense steady-state system A theta = b assembled in
// shared memory (fp64) and solved by Gaussian elimination; A is strictly diagonally dominant (diagonal
// -(1 or 2) - #neighbours, off-diagonals +1), so no pivoting is needed.  Site rate = theta on edge sites.
// Lattices above 160 sites (the dense matrix no longer fits shared memory; LSC_cat.py works for any `dimen`) use the
// structure of the system instead: M = -A is the typed-site graph Laplacian of the triangular lattice plus a positive
// diagonal -- symmetric, strictly diagonally dominant, condition number <= 14 -- so conjugate gradients on the lattice
// stencil (no matrix is stored) reach a 1e-15 relative residual in ~60 iterations; vectors in shared memory up to
// 76 x 76 sites, in an L2-resident global scratch above.
#include "so_internal.cuh"
#include <math.h>

namespace {

__device__ __constant__ int8_t a_ring1[6] = {1, 0, -1, -1, 0, 1};
__device__ __constant__ int8_t a_ring2[6] = {0, 1, 1, 0, -1, -1};

struct LscSmem {
  double *A;        // [N][N] row-major, only the leading ns x ns block is used
  double *b;        // [N]  right-hand side, then theta
  int *type;        // [N]  0 / 1 / 2
  int *idx;         // [N]  position among the typed sites, -1 if untyped
  int *sites;       // [N]  typed sites in ascending order
  uint32_t *bits;   // [W]
  int *ns;          // number of typed sites
};

__device__ __forceinline__ LscSmem carve(unsigned char *raw, int N, int W) {
  LscSmem s;
  s.A = (double *)raw;
  s.b = s.A + (size_t)N * N;
  s.type = (int *)(s.b + N);
  s.idx = s.type + N;
  s.sites = s.idx + N;
  s.bits = (uint32_t *)(s.sites + N);
  s.ns = (int *)(s.bits + W);
  return s;
}

// structure rate of the occupancy in s.bits; on return s.b[0..ns) = theta, s.type / s.sites / *s.ns valid.
// All threads of the block call this (contains __syncthreads()).
__device__ double lsc_solve(const LscSmem &s, int d, int N) {
  const int tid = threadIdx.x, nt = blockDim.x;
  BitView bv{s.bits, d, -1};
  for (int v = tid; v < N; v += nt) {
    int i2 = v / d, i1 = v - i2 * d;
    int x = bv.site(v);
    int ring = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) ring |= bv.at(i1 + a_ring1[k], i2 + a_ring2[k]) << k;
    int c6 = __popc(ring);
    int vac = (~ring) & 63;
    int adj = (vac & (((vac << 1) | (vac >> 5)) & 63)) != 0;
    s.type[v] = !x ? 0 : (c6 == 6 ? 1 : ((c6 == 4 && adj) ? 2 : 0));
  }
  __syncthreads();
  if (tid == 0) {
    int n = 0;
    for (int v = 0; v < N; ++v) {
      if (s.type[v]) { s.idx[v] = n; s.sites[n] = v; ++n; } else s.idx[v] = -1;
    }
    *s.ns = n;
  }
  __syncthreads();
  const int ns = *s.ns;
  if (ns == 0) return 0.0;
  for (int e = tid; e < ns * N; e += nt) s.A[e] = 0.0;     // rows of length N (lda = N)
  __syncthreads();
  for (int i = tid; i < ns; i += nt) {
    int v = s.sites[i];
    int i2 = v / d, i1 = v - i2 * d;
    double diag;
    if (s.type[v] == 1) { diag = -2.0; s.b[i] = -1.0; } else { diag = -1.0; s.b[i] = 0.0; }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      int w = wrap(i2 + a_ring2[k], d) * d + wrap(i1 + a_ring1[k], d);
      int j = s.idx[w];
      if (j >= 0) { diag -= 1.0; s.A[(size_t)i * N + j] += 1.0; }
    }
    s.A[(size_t)i * N + i] += diag;
  }
  __syncthreads();
  // Gaussian elimination, 16 x 16 thread tiling of the trailing block
  const int ty = tid >> 4, tx = tid & 15;
  for (int k = 0; k < ns - 1; ++k) {
    const double inv = 1.0 / s.A[(size_t)k * N + k];
    for (int i = k + 1 + tid; i < ns; i += nt) s.A[(size_t)i * N + k] *= inv;      // multipliers l_ik
    __syncthreads();
    for (int i = k + 1 + ty; i < ns; i += 16) {
      const double l = s.A[(size_t)i * N + k];
      if (l != 0.0) {
        for (int j = k + 1 + tx; j < ns; j += 16) s.A[(size_t)i * N + j] -= l * s.A[(size_t)k * N + j];
        if (tx == 0) s.b[i] -= l * s.b[k];
      }
    }
    __syncthreads();
  }
  // back substitution by warp 0
  if (tid < 32) {
    for (int k = ns - 1; k >= 0; --k) {
      double acc = 0.0;
      for (int j = k + 1 + tid; j < ns; j += 32) acc += s.A[(size_t)k * N + j] * s.b[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(SO_FULL, acc, o);
      if (tid == 0) s.b[k] = (s.b[k] - acc) / s.A[(size_t)k * N + k];
      __syncwarp();
    }
  }
  __syncthreads();
  // structure rate = sum of theta over edge sites (A_rate = k_des_edge on type 2), in ascending site order
  double rate = 0.0;
  for (int i = 0; i < ns; ++i)
    if (s.type[s.sites[i]] == 2) rate += s.b[i];
  return rate;
}

// ---- conjugate gradients on the lattice stencil (any lattice size) ---------------------------------------------------
struct CgMem {
  double *x, *r, *p, *q;   // [N] each, indexed by site (zero on untyped sites); shared or global
  uint8_t *type;           // [N] 0 / 1 / 2 (shared)
  uint32_t *bits;          // [W] (shared)
  double *red;             // [8] warp partials (shared)
};

__device__ __forceinline__ double block_sum(double v, double *red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SO_FULL, v, o);
  __syncthreads();                                 // red[] may still be read from the previous call
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];      // fixed order: deterministic
  return t;
}

// structure rate; on return m.x[v] = theta of typed site v (0 elsewhere), m.type valid.  Whole block.
__device__ double lsc_solve_cg(const CgMem &m, int d, int N) {
  const int tid = threadIdx.x, nt = blockDim.x;
  BitView bv{m.bits, d, -1};
  double rr = 0.0;
  for (int v = tid; v < N; v += nt) {
    int i2 = v / d, i1 = v - i2 * d;
    int x = bv.site(v);
    int ring = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) ring |= bv.at(i1 + a_ring1[k], i2 + a_ring2[k]) << k;
    int c6 = __popc(ring);
    int vac = (~ring) & 63;
    int adj = (vac & (((vac << 1) | (vac >> 5)) & 63)) != 0;
    int t = !x ? 0 : (c6 == 6 ? 1 : ((c6 == 4 && adj) ? 2 : 0));
    m.type[v] = (uint8_t)t;
    const double rhs = t == 1 ? 1.0 : 0.0;         // M theta = -b, b = -k_ads on terrace sites (LSC_cat.py:297)
    m.x[v] = 0.0; m.r[v] = rhs; m.p[v] = rhs;
    rr += rhs;                                     // rhs^2 = rhs
  }
  rr = block_sum(rr, m.red);
  const double stop = rr * 1e-30;                  // relative residual 1e-15
  for (int it = 0; it < 400 && rr > stop && rr > 0.0; ++it) {
    double pq = 0.0;
    for (int v = tid; v < N; v += nt) {
      const int t = m.type[v];
      double q = 0.0;
      if (t) {
        int i2 = v / d, i1 = v - i2 * d;
        double diag = t == 1 ? 2.0 : 1.0, nb = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          int w = wrap(i2 + a_ring2[k], d) * d + wrap(i1 + a_ring1[k], d);
          if (m.type[w]) { diag += 1.0; nb += m.p[w]; }
        }
        q = diag * m.p[v] - nb;
        pq += m.p[v] * q;
      }
      m.q[v] = q;
    }
    pq = block_sum(pq, m.red);
    const double alpha = rr / pq;
    double rr_new = 0.0;
    for (int v = tid; v < N; v += nt) {
      if (m.type[v]) {
        m.x[v] += alpha * m.p[v];
        const double r = m.r[v] - alpha * m.q[v];
        m.r[v] = r;
        rr_new += r * r;
      }
    }
    rr_new = block_sum(rr_new, m.red);
    const double beta = rr_new / rr;
    for (int v = tid; v < N; v += nt)
      if (m.type[v]) m.p[v] = m.r[v] + beta * m.p[v];
    rr = rr_new;
    __syncthreads();
  }
  double rate = 0.0;
  for (int v = tid; v < N; v += nt)
    if (m.type[v] == 2) rate += m.x[v];
  return block_sum(rate, m.red);
}

__device__ __forceinline__ CgMem carve_cg(unsigned char *raw, int N, int W, double *scratch) {
  CgMem m;
  double *base = scratch ? scratch : (double *)raw;
  m.x = base; m.r = base + N; m.p = base + 2 * (size_t)N; m.q = base + 3 * (size_t)N;
  unsigned char *rest = scratch ? raw : raw + (size_t)4 * N * 8;
  m.red = (double *)rest;
  m.bits = (uint32_t *)(m.red + 8);
  m.type = (uint8_t *)(m.bits + W);
  return m;
}

__global__ void __launch_bounds__(256) lsc_analytical_cg_kernel(const uint32_t *__restrict__ bits, long long n, int d, int N,
                                                                int W, double *site_rates, double *struct_rate,
                                                                double *scratch) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  CgMem m = carve_cg(s_raw, N, W, scratch ? scratch + (size_t)blockIdx.x * 4 * N : nullptr);
  for (long long c = blockIdx.x; c < n; c += gridDim.x) {
    __syncthreads();
    for (int w = threadIdx.x; w < W; w += blockDim.x) m.bits[w] = bits[c * W + w];
    __syncthreads();
    double rate = lsc_solve_cg(m, d, N);
    if (site_rates)
      for (int v = threadIdx.x; v < N; v += blockDim.x) site_rates[c * N + v] = m.type[v] == 2 ? m.x[v] : 0.0;
    if (struct_rate && threadIdx.x == 0) struct_rate[c] = rate;
  }
}

__global__ void __launch_bounds__(256) sa_analytical_cg_kernel(SaParams P, double *of_out, double *scratch) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  CgMem m = carve_cg(s_raw, P.N, P.W, scratch ? scratch + (size_t)blockIdx.x * 4 * P.N : nullptr);
  __shared__ int s_accept;
  for (long long c = blockIdx.x; c < P.n_chains; c += gridDim.x) {
    __syncthreads();
    uint32_t *bits_g = P.bits + c * P.W;
    for (int w = threadIdx.x; w < P.W; w += blockDim.x) m.bits[w] = bits_g[w];
    __syncthreads();
    double OF = lsc_solve_cg(m, P.d, P.N);
    const double tsc = P.tscale[c];
    unsigned n_acc = 0;
    for (long long st = 0; st < P.n_steps; ++st) {
      int f;
      float u;
      if (P.proposals) {
        f = __ldg(P.proposals + c * P.n_steps + st);
        u = __ldg(P.uniforms + c * P.n_steps + st);
      } else {
        unsigned long long step = (unsigned long long)(P.step0 + st), cid = (unsigned long long)(P.chain_id0 + c);
        uint32_t o[4];
        philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), (uint32_t)cid, (uint32_t)(cid >> 32), (uint32_t)P.seed,
                      (uint32_t)(P.seed >> 32), o);
        f = (int)__umulhi(o[0], (uint32_t)P.N);
        u = (float)(o[1] >> 8) * 5.9604644775390625e-08f;
      }
      __syncthreads();
      if (threadIdx.x == 0) m.bits[f >> 5] ^= (1u << (f & 31));
      __syncthreads();
      const double OF_new = lsc_solve_cg(m, P.d, P.N);
      if (threadIdx.x == 0) {
        const double T = __dmul_rn(tsc, __ldg(P.temps + st));
        const double delta = OF_new - OF;
        bool acc;
        if (delta > 0.0) acc = true;
        else if (T > 0.0) acc = exp(delta / T) > (double)u;
        else acc = false;
        s_accept = acc ? 1 : 0;
        if (!acc) m.bits[f >> 5] ^= (1u << (f & 31));
        if (P.accept_out) P.accept_out[c * P.n_steps + st] = acc ? 1 : 0;
      }
      __syncthreads();
      if (s_accept) { OF = OF_new; ++n_acc; }
    }
    __syncthreads();
    for (int w = threadIdx.x; w < P.W; w += blockDim.x) bits_g[w] = m.bits[w];
    if (threadIdx.x == 0) {
      if (of_out) of_out[c] = OF;
      if (n_acc) atomicAdd(P.counters + 1, (unsigned long long)n_acc);
    }
  }
}

__global__ void __launch_bounds__(256) lsc_analytical_kernel(const uint32_t *__restrict__ bits, long long n, int d, int N,
                                                             int W, double *site_rates, double *struct_rate) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  LscSmem s = carve(s_raw, N, W);
  for (long long c = blockIdx.x; c < n; c += gridDim.x) {
    __syncthreads();
    for (int w = threadIdx.x; w < W; w += blockDim.x) s.bits[w] = bits[c * W + w];
    __syncthreads();
    double rate = lsc_solve(s, d, N);
    const int ns = *s.ns;
    if (site_rates) {
      for (int v = threadIdx.x; v < N; v += blockDim.x) site_rates[c * N + v] = 0.0;
      __syncthreads();
      for (int i = threadIdx.x; i < ns; i += blockDim.x)
        if (s.type[s.sites[i]] == 2) site_rates[c * N + s.sites[i]] = s.b[i];
    }
    if (struct_rate && threadIdx.x == 0) struct_rate[c] = rate;
  }
}

__global__ void __launch_bounds__(256) sa_analytical_kernel(SaParams P, double *of_out) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  LscSmem s = carve(s_raw, P.N, P.W);
  __shared__ int s_accept;
  for (long long c = blockIdx.x; c < P.n_chains; c += gridDim.x) {
    __syncthreads();
    uint32_t *bits_g = P.bits + c * P.W;
    for (int w = threadIdx.x; w < P.W; w += blockDim.x) s.bits[w] = bits_g[w];
    __syncthreads();
    double OF = lsc_solve(s, P.d, P.N);
    const double tsc = P.tscale[c];
    unsigned n_acc = 0;
    for (long long st = 0; st < P.n_steps; ++st) {
      int f;
      float u;
      if (P.proposals) {
        f = __ldg(P.proposals + c * P.n_steps + st);
        u = __ldg(P.uniforms + c * P.n_steps + st);
      } else {
        unsigned long long step = (unsigned long long)(P.step0 + st), cid = (unsigned long long)(P.chain_id0 + c);
        uint32_t o[4];
        philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), (uint32_t)cid, (uint32_t)(cid >> 32), (uint32_t)P.seed,
                      (uint32_t)(P.seed >> 32), o);
        f = (int)__umulhi(o[0], (uint32_t)P.N);
        u = (float)(o[1] >> 8) * 5.9604644775390625e-08f;
      }
      __syncthreads();
      if (threadIdx.x == 0) s.bits[f >> 5] ^= (1u << (f & 31));
      __syncthreads();
      const double OF_new = lsc_solve(s, P.d, P.N);
      if (threadIdx.x == 0) {
        const double T = __dmul_rn(tsc, __ldg(P.temps + st));
        const double delta = OF_new - OF;
        bool acc;
        if (delta > 0.0) acc = true;
        else if (T > 0.0) acc = exp(delta / T) > (double)u;
        else acc = false;
        s_accept = acc ? 1 : 0;
        if (!acc) s.bits[f >> 5] ^= (1u << (f & 31));
        if (P.accept_out) P.accept_out[c * P.n_steps + st] = acc ? 1 : 0;
      }
      __syncthreads();
      if (s_accept) { OF = OF_new; ++n_acc; }
    }
    __syncthreads();
    for (int w = threadIdx.x; w < P.W; w += blockDim.x) bits_g[w] = s.bits[w];
    if (threadIdx.x == 0) {
      if (of_out) of_out[c] = OF;
      if (n_acc) atomicAdd(P.counters + 1, (unsigned long long)n_acc);
    }
  }
}

size_t lsc_smem_bytes(int N, int W) { return (size_t)N * N * 8 + (size_t)N * 8 + (size_t)3 * N * 4 + (size_t)W * 4 + 16; }

}  // namespace

int so_lsc_max_sites(void) { return 1024 * 1024; }   // any lattice so_lattice_create accepts
static const int kDenseMaxSites = 160;               // N*N fp64 must fit the 227 KB of shared memory of one CTA

// shared memory of the CG kernels: 4 vectors when they fit (<= 200 KB), else only the small arrays
static size_t cg_smem_bytes(int N, int W, bool *vectors_in_smem) {
  size_t small = 64 + (size_t)W * 4 + (size_t)N + 16;
  size_t vec = (size_t)4 * N * 8;
  *vectors_in_smem = vec + small <= 200 * 1024;
  return small + (*vectors_in_smem ? vec : 0);
}

int so_launch_lsc_analytical(const so_lattice *lat, const uint32_t *bits_dev, int64_t n, double *site_rates_dev,
                             double *struct_rate_dev, cudaStream_t st) {
  if (n == 0) return SO_OK;
  int n_sm = 0;
  SO_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, lat->device));
  int blocks = (int)(n < (int64_t)n_sm * 4 ? n : (int64_t)n_sm * 4);
  if (lat->N > kDenseMaxSites) {
    bool in_smem;
    size_t smem = cg_smem_bytes(lat->N, lat->W, &in_smem);
    double *scratch = nullptr;
    if (!in_smem) SO_CUDA(cudaMallocAsync((void **)&scratch, (size_t)blocks * 4 * lat->N * 8, st));
    SO_CUDA(cudaFuncSetAttribute(lsc_analytical_cg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lsc_analytical_cg_kernel<<<blocks, 256, smem, st>>>(bits_dev, n, lat->d, lat->N, lat->W, site_rates_dev,
                                                        struct_rate_dev, scratch);
    SO_CUDA(cudaGetLastError());
    if (scratch) SO_CUDA(cudaFreeAsync(scratch, st));
    return SO_OK;
  }
  size_t smem = lsc_smem_bytes(lat->N, lat->W);
  SO_CUDA(cudaFuncSetAttribute(lsc_analytical_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  lsc_analytical_kernel<<<blocks, 256, smem, st>>>(bits_dev, n, lat->d, lat->N, lat->W, site_rates_dev, struct_rate_dev);
  SO_CUDA(cudaGetLastError());
  return SO_OK;
}

int so_launch_anneal_analytical(const so_chains *c, const SaParams &P, double *of_dev, cudaStream_t st) {
  int n_sm = 0;
  SO_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, c->device));
  int blocks = (int)(P.n_chains < (long long)n_sm * 4 ? P.n_chains : (long long)n_sm * 4);
  if (P.N > kDenseMaxSites) {
    bool in_smem;
    size_t smem = cg_smem_bytes(P.N, P.W, &in_smem);
    double *scratch = nullptr;
    if (!in_smem) SO_CUDA(cudaMallocAsync((void **)&scratch, (size_t)blocks * 4 * P.N * 8, st));
    SO_CUDA(cudaFuncSetAttribute(sa_analytical_cg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sa_analytical_cg_kernel<<<blocks, 256, smem, st>>>(P, of_dev, scratch);
    SO_CUDA(cudaGetLastError());
    if (scratch) SO_CUDA(cudaFreeAsync(scratch, st));
    return SO_OK;
  }
  size_t smem = lsc_smem_bytes(P.N, P.W);
  SO_CUDA(cudaFuncSetAttribute(sa_analytical_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sa_analytical_kernel<<<blocks, 256, smem, st>>>(P, of_dev);
  SO_CUDA(cudaGetLastError());
  return SO_OK;
}
