// so_train.cu -- the surrogate's regressor trained on the device: the K -> H -> 1 ReLU MLP of
// OML/train_surrogate.py:141-152, mini-batch Adam with L2 penalty exactly as scikit-learn's MLPRegressor runs it
// (squared loss / 2, alpha / 2 |coefs|^2 / batch, one Adam step per batch, tol-based stop on the epoch loss), in fp64.
//
// Training is a strictly sequential chain of tiny batches (200 rows x K x H: ~0.6 M fp64 additions per step after
// exploiting the 0/1 inputs, up to ~1e6 steps), so it runs as ONE persistent thread-block cluster of 8 CTAs with the
// whole optimiser state in (distributed) shared memory: every CTA keeps a replica of the parameters and works on an
// eighth of each batch's rows (rows as bit-vectors, expanded to 0/1 operands of fp64 tensor-core tiles for the two small
// GEMMs of a step); the gradients are reduce-scattered through
// DSMEM, each CTA applies Adam to the parameter slice whose moments it owns and stores the result into all replicas.
// Nothing but the batch rows (K/8 B each) and the targets is read from HBM per step.  The order of the samples in every
// epoch is supplied by the caller, so a run can replay scikit-learn's own shuffles (parity tests) or any other stream.
#include "so_internal.cuh"
#include <cooperative_groups.h>
#include <math.h>
#include <string.h>

namespace cg = cooperative_groups;

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

struct TrainState {        // persists in global memory between launches
  long long t;             // Adam steps so far
  double best;             // best epoch loss
  int count;               // epochs without improvement by tol
  int n_iter;              // epochs run
  int stopped;
};

struct so_trainer {
  int device, K, KW, H, batch, n_iter_no_change;
  long long n_rows;
  double lr, alpha, beta1, beta2, eps, tol;
  uint32_t *xbits_dev;     // [n_rows][KW]
  double *y_dev;           // [n_rows]
  double *theta_dev;       // [3][P]: parameters, first and second Adam moment; P = (K+1)*H + H + 1
  TrainState *state_dev;
  cudaStream_t stream;
  cudaEvent_t ev0, ev1;
  size_t smem;
};

namespace {

constexpr int TRAIN_THREADS = 1024;
constexpr int TRAIN_CTAS = 8;            // one thread-block cluster
constexpr int TRAIN_MAX_BATCH = 1024;    // <= 128 rows per CTA (four row words)

// D (8x8) += A (8x4, row-major) * B (4x8, column-major) on the fp64 tensor cores.  Fragments (lane l): a = A[l/4][l%4],
// b = B[l%4][l/4], c0/c1 = C[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ double bit_to_double(unsigned bit) { return __hiloint2double((int)(bit * 0x3ff00000u), 0); }

struct TrainLayout {       // shared memory of one CTA, in doubles unless noted
  int P1, P, slice0, slice1, rows_max, RW;
  int KS, Hpad;
  size_t th, gp, mo, ve, act, del, d2, sqr, misc, fpart, xb, bytes;
};
__host__ __device__ inline TrainLayout train_layout(int K, int H, int batch, int rank) {
  TrainLayout L;
  L.P1 = (K + 1) * H;
  L.P = L.P1 + H + 1;
  L.slice0 = (int)((long long)L.P * rank / TRAIN_CTAS);
  L.slice1 = (int)((long long)L.P * (rank + 1) / TRAIN_CTAS);
  L.rows_max = (batch + TRAIN_CTAS - 1) / TRAIN_CTAS;
  L.RW = (L.rows_max + 31) / 32;
  const int KW = (K + 31) / 32;
  const int slice_max = L.P / TRAIN_CTAS + 1;
  size_t o = 0;
  L.th = o; o += L.P;                        // all parameters (replica)
  L.gp = o; o += L.P;                        // this CTA's partial gradients; [P] = its sum of squared residuals,
  o += 2;                                    // [P + 1] = sum of squares of the coefficients in its slice
  L.mo = o; o += slice_max;                  // Adam moments of this CTA's parameter slice
  L.ve = o; o += slice_max;
  L.act = o; o += (size_t)L.rows_max * H;    // activations of this CTA's rows
  L.del = o; o += (size_t)L.rows_max * H;    // delta1 of this CTA's rows
  L.d2 = o; o += L.rows_max;                 // targets, then residuals
  L.sqr = o; o += L.rows_max;                // squared residuals
  L.misc = o; o += 72;                       // block reduction scratch [0..32], [36], [40..71]; lr_t [33], [35]; stop flag [34]
  // forward tiles: MT x NT output tiles of 8 x 8, the K dimension cut into KS slices so that all 32 warps have a tile
  L.Hpad = (H + 7) & ~7;
  {
    const int tiles = ((L.rows_max + 7) / 8) * (L.Hpad / 8);
    L.KS = tiles >= 32 ? 1 : (32 / tiles > 4 ? 4 : 32 / tiles);
  }
  L.fpart = o; o += (size_t)L.KS * (((L.rows_max + 7) / 8) * 8) * L.Hpad;   // [KS][rows padded to 8][Hpad] partial pre-activations
  L.xb = o * 8; o += ((size_t)L.rows_max * KW + 1) / 2;      // u32 [rows][KW]
  L.bytes = o * 8 + 16;
  return L;
}

// One cluster of 8 CTAs runs the whole optimisation.  Each CTA holds a replica of the parameters, takes 1/8 of the
// rows of every batch (forward, residuals, partial gradients over its rows), then the cluster reduce-scatters the
// gradients through distributed shared memory: CTA c sums the eight partials of its parameter slice in rank order,
// applies Adam with the moments it owns and stores the new values into all eight replicas.  Two cluster barriers per
// batch.  parameter vector: [ (K+1)*H : W1 rows 0..K-1, then b1 as row K ][ H : w2 ][ 1 : b2 ]
__global__ void __cluster_dims__(TRAIN_CTAS, 1, 1) __launch_bounds__(TRAIN_THREADS, 1)
    train_kernel(int K, int KW, int H, int batch, long long n_rows, const uint32_t *__restrict__ xbits,
                 const double *__restrict__ y, double *theta_g, TrainState *state_g, int n_epochs,
                 const int32_t *__restrict__ orders, double *loss_out, double lr, double alpha, double beta1, double beta2,
                 double eps, double tol, int n_iter_no_change) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const TrainLayout L = train_layout(K, H, batch, rank);
  const int P1 = L.P1, P = L.P;
  double *base = (double *)s_raw;
  double *th = base + L.th, *gp = base + L.gp, *mo = base + L.mo, *ve = base + L.ve, *act = base + L.act;
  double *d2 = base + L.d2, *misc = base + L.misc, *del = base + L.del, *sqr = base + L.sqr;
  uint32_t *xb = (uint32_t *)(s_raw + L.xb);
  double *fpart = base + L.fpart;
  __shared__ TrainState st;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const double *gp_of[TRAIN_CTAS];
  double *th_of[TRAIN_CTAS];
#pragma unroll
  for (int c = 0; c < TRAIN_CTAS; ++c) {
    gp_of[c] = cluster.map_shared_rank(gp, c);
    th_of[c] = cluster.map_shared_rank(th, c);
  }

  for (int i = tid; i < P; i += blockDim.x) th[i] = theta_g[i];
  for (int i = L.slice0 + tid; i < L.slice1; i += blockDim.x) {
    mo[i - L.slice0] = theta_g[P + i];
    ve[i - L.slice0] = theta_g[2 * P + i];
  }
  if (tid == 0) { st = *state_g; misc[34] = st.stopped ? 1.0 : 0.0; }
  __syncthreads();
  cluster.sync();

  // the rows of the NEXT batch travel in registers while the current one is processed (order -> row: two dependent
  // global round trips per batch otherwise)
  constexpr int EPT = 4;
  const bool prefetch = L.rows_max * KW <= EPT * (int)blockDim.x;
  uint32_t px[EPT];
  double py = 0.0;
  auto fetch = [&](int e2, long long s02) {
    const int32_t *order2 = orders + (size_t)e2 * n_rows;
    const int nb2 = (int)min((long long)batch, n_rows - s02);
    const int lo2 = (int)((long long)nb2 * rank / TRAIN_CTAS), nr2 = (int)((long long)nb2 * (rank + 1) / TRAIN_CTAS) - lo2;
#pragma unroll
    for (int u = 0; u < EPT; ++u) {
      const int i = tid + u * (int)blockDim.x;
      if (i < nr2 * KW) {
        const int r = i / KW, w = i - r * KW;
        px[u] = xbits[(size_t)order2[s02 + lo2 + r] * KW + w];
      }
    }
    if (tid < nr2) py = y[order2[s02 + lo2 + tid]];
  };
  // beta^t as running products (pow() once per launch: a per-step pow on one thread stalled every barrier behind it)
  long long t = st.t;
  double b1t = pow(beta1, (double)t), b2t = pow(beta2, (double)t);
  if (prefetch && n_epochs > 0 && misc[34] == 0.0) fetch(0, 0);

  for (int e = 0; e < n_epochs && misc[34] == 0.0; ++e) {
    const int32_t *order = orders + (size_t)e * n_rows;
    double acc_loss = 0.0;                 // rank 0, thread 0 only
    for (long long s0 = 0; s0 < n_rows; s0 += batch) {
      const int nb = (int)min((long long)batch, n_rows - s0);
      const int r_lo = (int)((long long)nb * rank / TRAIN_CTAS), r_hi = (int)((long long)nb * (rank + 1) / TRAIN_CTAS);
      const int nr = r_hi - r_lo;          // this CTA's rows of the batch
      t += 1;
      b1t *= beta1;
      b2t *= beta2;
      // ---- rows (bit-vectors) and targets ----
      if (prefetch) {
#pragma unroll
        for (int u = 0; u < EPT; ++u) {
          const int i = tid + u * (int)blockDim.x;
          if (i < nr * KW) xb[i] = px[u];
        }
        if (tid < nr) d2[tid] = py;
      } else {
        for (int i = tid; i < nr * KW; i += blockDim.x) {
          const int r = i / KW, w = i - r * KW;
          xb[i] = xbits[(size_t)order[s0 + r_lo + r] * KW + w];
        }
        if (tid < nr) d2[tid] = y[order[s0 + r_lo + tid]];
      }
      __syncthreads();
      if (prefetch) {
        if (s0 + batch < n_rows) fetch(e, s0 + batch);
        else if (e + 1 < n_epochs) fetch(e + 1, 0);
      }
      // ---- forward, first layer on the fp64 tensor cores: X (rows x K, 0/1 from the bit-vectors) * W1 (K x H) ----
      {
        const int MT = (nr + 7) >> 3, NT = L.Hpad >> 3, KS = L.KS, rows_pad = ((L.rows_max + 7) >> 3) << 3;
        const int ksteps = (K + 3) >> 2, per = (ksteps + KS - 1) / KS;
        const int ar = lane >> 2, ak = lane & 3;                  // fragment coordinates of this lane
        for (int t = warp; t < MT * NT * KS; t += (int)(blockDim.x >> 5)) {
          const int ks = t / (MT * NT), mt = (t / NT) % MT, nt = t % NT;
          const int row = 8 * mt + ar, col = 8 * nt + ar;
          const uint32_t *xrow = xb + (row < nr ? row : 0) * KW;
          double c0 = 0.0, c1 = 0.0;
          const int kk_hi = min(ksteps, (ks + 1) * per);
          for (int kk = ks * per; kk < kk_hi; ++kk) {
            const int k = 4 * kk + ak;
            const unsigned bit = (row < nr && k < K) ? (xrow[k >> 5] >> (k & 31)) & 1u : 0u;
            const double b = (k < K && col < H) ? th[k * H + col] : 0.0;
            dmma884(c0, c1, bit_to_double(bit), b);
          }
          double *o = fpart + ((size_t)ks * rows_pad + 8 * mt + ar) * L.Hpad + 8 * nt + 2 * ak;
          o[0] = c0;
          o[1] = c1;
        }
      }
      __syncthreads();
      // ---- one warp per row: bias + relu, output, residual, delta1 (lane j = hidden unit j; no barrier in between) ----
      {
        const int rows_pad = ((L.rows_max + 7) >> 3) << 3;
        for (int r = warp; r < nr; r += (int)(blockDim.x >> 5)) {
          double a = 0.0, w2j = 0.0;
          if (lane < H) {
            double h = th[K * H + lane];
            for (int ks = 0; ks < L.KS; ++ks) h += fpart[((size_t)ks * rows_pad + r) * L.Hpad + lane];   // fixed order
            a = fmax(h, 0.0);
            w2j = th[P1 + lane];
            act[r * H + lane] = a;
          }
          double pred = a * w2j;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) pred += __shfl_xor_sync(SO_FULL, pred, o);
          const double dd = pred + th[P1 + H] - d2[r];
          if (lane < H) del[r * H + lane] = a > 0.0 ? dd * w2j : 0.0;
          __syncwarp();
          if (lane == 0) { d2[r] = dd; sqr[r] = dd * dd; }
        }
      }
      __syncthreads();
      // ---- output-layer partial gradients (warp j < H: column j of a^T delta2; warp H: sum of delta2) and the loss terms
      //      (last warp: squared residuals of this CTA's rows, squared coefficients of its slice) ----
      if (warp <= H) {
        double g = 0.0;
        for (int r = lane; r < nr; r += 32) g += (warp < H ? act[r * H + warp] : 1.0) * d2[r];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(SO_FULL, g, o);
        if (lane == 0) gp[P1 + warp] = g;
      }
      if (warp == (int)(blockDim.x >> 5) - 1) {
        double part = 0.0, part2 = 0.0;
        for (int r = lane; r < nr; r += 32) part += sqr[r];
        for (int p = L.slice0 + lane; p < L.slice1; p += 32)
          if (p < K * H || (p >= P1 && p < P1 + H)) part2 += th[p] * th[p];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          part += __shfl_xor_sync(SO_FULL, part, o);
          part2 += __shfl_xor_sync(SO_FULL, part2, o);
        }
        if (lane == 0) { gp[P] = part; gp[P + 1] = part2; }
      }
      // ---- partial first-layer gradients on the fp64 tensor cores: X^T ((K+1) x rows; row K = ones: the bias) * delta1 ----
      {
        const int MT = (K + 1 + 7) >> 3, NT = L.Hpad >> 3, rsteps = (nr + 3) >> 2;
        const int ar = lane >> 2, ak = lane & 3;
        for (int t = warp; t < MT * NT; t += (int)(blockDim.x >> 5)) {
          const int mt = t / NT, nt = t - mt * NT;
          const int k = 8 * mt + ar, col = 8 * nt + ar;
          double c0 = 0.0, c1 = 0.0;
          for (int kk = 0; kk < rsteps; ++kk) {
            const int r = 4 * kk + ak;
            const unsigned bit = r < nr ? (k < K ? (xb[r * KW + (k >> 5)] >> (k & 31)) & 1u : (k == K ? 1u : 0u)) : 0u;
            const double b = (r < nr && col < H) ? del[r * H + col] : 0.0;
            dmma884(c0, c1, bit_to_double(bit), b);
          }
          const int oc = 8 * nt + 2 * ak;
          if (k <= K) {
            if (oc < H) gp[k * H + oc] = c0;
            if (oc + 1 < H) gp[k * H + oc + 1] = c1;
          }
        }
      }
      cluster.sync();                      // all partial gradients are in place
      // ---- reduce-scatter + Adam on this CTA's slice; new values go to every replica ----
      const double lr_t = lr * sqrt(1.0 - b2t) / (1.0 - b1t);
      for (int p = L.slice0 + tid; p < L.slice1; p += blockDim.x) {
        double g = 0.0;
#pragma unroll
        for (int c = 0; c < TRAIN_CTAS; ++c) g += gp_of[c][p];
        if (p < K * H || (p >= P1 && p < P1 + H)) g += alpha * th[p];      // L2 penalty on the coefficients only
        g /= (double)nb;
        const int q = p - L.slice0;
        const double m = beta1 * mo[q] + (1.0 - beta1) * g;
        const double v = beta2 * ve[q] + (1.0 - beta2) * (g * g);
        mo[q] = m;
        ve[q] = v;
        const double nv = th[p] + -lr_t * m / (sqrt(v) + eps);
#pragma unroll
        for (int c = 0; c < TRAIN_CTAS; ++c) th_of[c][p] = nv;
      }
      if (rank == 0 && tid == 0) {
        double sq_all = 0.0, wsq = 0.0;
#pragma unroll
        for (int c = 0; c < TRAIN_CTAS; ++c) {
          sq_all += gp_of[c][P];
          wsq += gp_of[c][P + 1];
        }
        acc_loss += (0.5 * sq_all / (double)nb + 0.5 * alpha * wsq / (double)nb) * nb;
      }
      cluster.sync();                      // every replica holds the new parameters; partials may be overwritten
    }
    if (rank == 0 && tid == 0) {
      const double cur = acc_loss / (double)n_rows;
      loss_out[e] = cur;
      st.n_iter += 1;
      st.count = cur > st.best - tol ? st.count + 1 : 0;
      if (cur < st.best) st.best = cur;
      if (st.count > n_iter_no_change) {
        st.stopped = 1;
        for (int c = 0; c < TRAIN_CTAS; ++c) cluster.map_shared_rank(misc, c)[34] = 1.0;
      }
    }
    cluster.sync();
  }
  for (int i = L.slice0 + tid; i < L.slice1; i += blockDim.x) {
    theta_g[i] = th[i];
    theta_g[P + i] = mo[i - L.slice0];
    theta_g[2 * P + i] = ve[i - L.slice0];
  }
  if (rank == 0 && tid == 0) { st.t = t; *state_g = st; }
  cluster.sync();                          // no CTA leaves while its shared memory may still be addressed
}

// forward pass of explicit 0/1 rows with the current parameters (MLPRegressor.predict): one warp per row
__global__ void predict_kernel(int K, int H, long long n, const uint8_t *__restrict__ X, const double *__restrict__ theta,
                               double *out) {
  const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n) return;
  const int P1 = (K + 1) * H;
  double term = 0.0;
  if (lane < H) {
    double h = theta[K * H + lane];
    for (int k = 0; k < K; ++k)
      if (X[r * K + k]) h += theta[k * H + lane];
    term = fmax(h, 0.0) * theta[P1 + lane];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(SO_FULL, term, o);
  if (lane == 0) out[r] = term + theta[P1 + H];
}

size_t train_smem_bytes(int K, int H, int batch) { return train_layout(K, H, batch, 0).bytes; }

}  // namespace

extern "C" {

int so_trainer_create(int device, int64_t n_rows, int K, int H, const uint8_t *X, const double *y, const double *W1,
                      const double *b1, const double *w2, double b2, const so_train_params *hp, so_trainer **out) {
  SO_REQUIRE(out, "out is NULL");
  *out = nullptr;
  SO_REQUIRE(X && y && W1 && b1 && w2 && hp, "NULL argument");
  SO_REQUIRE(n_rows >= 1 && n_rows < (1LL << 31), "n_rows=%lld out of range", (long long)n_rows);
  SO_REQUIRE(K >= 1 && H >= 1 && H <= 31, "K=%d, H=%d: need K >= 1 and 1 <= H <= 31", K, H);
  SO_REQUIRE(hp->batch_size >= 1 && hp->batch_size <= TRAIN_MAX_BATCH, "batch_size=%d out of range [1,%d]", hp->batch_size,
             TRAIN_MAX_BATCH);
  SO_REQUIRE(hp->lr > 0 && hp->alpha >= 0 && hp->beta1 >= 0 && hp->beta1 < 1 && hp->beta2 >= 0 && hp->beta2 < 1 && hp->eps > 0,
             "invalid Adam hyper-parameters");
  int n_dev = 0;
  SO_CUDA(cudaGetDeviceCount(&n_dev));
  SO_REQUIRE(device >= 0 && device < n_dev, "device %d out of range (%d visible)", device, n_dev);
  SO_CUDA(cudaSetDevice(device));
  const int batch = (int)std::min<int64_t>(hp->batch_size, n_rows);
  const size_t smem = train_smem_bytes(K, H, batch);
  SO_REQUIRE(smem <= 227 * 1024, "descriptor too wide for the on-chip trainer: K=%d, H=%d, batch=%d need %zu B of shared memory",
             K, H, batch, smem);
  so_trainer *t = new (std::nothrow) so_trainer();
  if (!t) { so_set_error("out of host memory"); return SO_ENOMEM; }
  memset(t, 0, sizeof(*t));
  t->device = device; t->K = K; t->KW = (K + 31) / 32; t->H = H; t->batch = batch; t->n_rows = n_rows;
  t->lr = hp->lr; t->alpha = hp->alpha; t->beta1 = hp->beta1; t->beta2 = hp->beta2; t->eps = hp->eps; t->tol = hp->tol;
  t->n_iter_no_change = hp->n_iter_no_change; t->smem = smem;
  const size_t P = (size_t)(K + 1) * H + H + 1;
  cudaError_t e = cudaMalloc(&t->xbits_dev, (size_t)n_rows * t->KW * 4);
  if (e == cudaSuccess) e = cudaMalloc(&t->y_dev, (size_t)n_rows * 8);
  if (e == cudaSuccess) e = cudaMalloc(&t->theta_dev, 3 * P * 8);
  if (e == cudaSuccess) e = cudaMalloc(&t->state_dev, sizeof(TrainState));
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&t->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&t->ev1);
  if (e != cudaSuccess) {
    so_set_error("trainer allocation failed: %s", cudaGetErrorString(e));
    so_trainer_destroy(t);
    return SO_ENOMEM;
  }
  // rows -> bit-vectors (the occupancy pack kernel: K plays the role of the site count); non-0/1 inputs are refused
  uint8_t *x_dev = nullptr;
  unsigned long long *flag = nullptr;
  const int64_t slab = std::max<int64_t>(1, (256LL << 20) / K);
  int rc = SO_OK;
  if (cudaMalloc(&x_dev, (size_t)std::min(slab, n_rows) * K) != cudaSuccess || cudaMalloc(&flag, 8) != cudaSuccess) {
    so_set_error("out of device memory staging the training rows");
    rc = SO_ENOMEM;
  }
  unsigned long long bad = 0;
  if (rc == SO_OK) {
    cudaMemsetAsync(flag, 0, 8, t->stream);
    for (int64_t r0 = 0; r0 < n_rows && rc == SO_OK; r0 += slab) {
      const int64_t nr = std::min(slab, n_rows - r0);
      if (cudaMemcpyAsync(x_dev, X + r0 * K, (size_t)nr * K, cudaMemcpyHostToDevice, t->stream) != cudaSuccess) rc = SO_ECUDA;
      if (rc == SO_OK) rc = so_launch_pack(x_dev, t->xbits_dev + r0 * t->KW, nr, K, t->KW, flag, t->stream);
      if (rc == SO_OK && cudaStreamSynchronize(t->stream) != cudaSuccess) rc = SO_ECUDA;
    }
    if (rc == SO_OK && cudaMemcpy(&bad, flag, 8, cudaMemcpyDeviceToHost) != cudaSuccess) rc = SO_ECUDA;
    if (rc == SO_ECUDA) so_set_error("CUDA error while staging the training rows: %s", cudaGetErrorString(cudaGetLastError()));
  }
  cudaFree(x_dev);
  cudaFree(flag);
  if (rc == SO_OK && bad) { so_set_error("training rows must be 0/1 occupancies"); rc = SO_EINVAL; }
  if (rc == SO_OK) {
    std::vector<double> theta(3 * P, 0.0);
    for (int k = 0; k < K; ++k)
      for (int j = 0; j < H; ++j) theta[(size_t)k * H + j] = W1[(size_t)k * H + j];
    for (int j = 0; j < H; ++j) { theta[(size_t)K * H + j] = b1[j]; theta[(size_t)(K + 1) * H + j] = w2[j]; }
    theta[P - 1] = b2;
    TrainState st;
    st.t = 0; st.best = INFINITY; st.count = 0; st.n_iter = 0; st.stopped = 0;
    if (cudaMemcpy(t->theta_dev, theta.data(), 3 * P * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(t->y_dev, y, (size_t)n_rows * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(t->state_dev, &st, sizeof(st), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaFuncSetAttribute(train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      so_set_error("CUDA error while initialising the trainer: %s", cudaGetErrorString(cudaGetLastError()));
      rc = SO_ECUDA;
    }
  }
  if (rc != SO_OK) { so_trainer_destroy(t); return rc; }
  *out = t;
  return SO_OK;
}

int so_trainer_destroy(so_trainer *t) {
  if (!t) return SO_OK;
  cudaSetDevice(t->device);
  if (t->stream) cudaStreamSynchronize(t->stream);
  cudaFree(t->xbits_dev); cudaFree(t->y_dev); cudaFree(t->theta_dev); cudaFree(t->state_dev);
  if (t->ev0) cudaEventDestroy(t->ev0);
  if (t->ev1) cudaEventDestroy(t->ev1);
  if (t->stream) cudaStreamDestroy(t->stream);
  delete t;
  return SO_OK;
}

int so_trainer_run(so_trainer *t, int n_epochs, const int32_t *orders, double *loss_out, so_train_status *status) {
  SO_REQUIRE(t, "NULL trainer");
  SO_REQUIRE(n_epochs >= 0, "n_epochs=%d is negative", n_epochs);
  SO_REQUIRE((orders && loss_out) || n_epochs == 0, "orders / loss_out is NULL");
  SO_CUDA(cudaSetDevice(t->device));
  float ms = 0.f;
  TrainState before, after;
  SO_CUDA(cudaMemcpy(&before, t->state_dev, sizeof(before), cudaMemcpyDeviceToHost));
  after = before;
  if (n_epochs > 0 && !before.stopped) {
    const size_t n = (size_t)n_epochs * (size_t)t->n_rows;
    for (size_t i = 0; i < n; ++i)
      SO_REQUIRE(orders[i] >= 0 && orders[i] < t->n_rows, "orders[%zu] = %d is not a row index", i, orders[i]);
    int32_t *ord_dev = nullptr;
    double *loss_dev = nullptr;
    if (cudaMalloc(&ord_dev, n * 4) != cudaSuccess || cudaMalloc(&loss_dev, (size_t)n_epochs * 8) != cudaSuccess) {
      cudaFree(ord_dev);
      cudaGetLastError();
      so_set_error("out of device memory for %d epoch orders", n_epochs);
      return SO_ENOMEM;
    }
    cudaError_t e = cudaMemcpyAsync(ord_dev, orders, n * 4, cudaMemcpyHostToDevice, t->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(loss_dev, 0, (size_t)n_epochs * 8, t->stream);
    if (e == cudaSuccess) e = cudaEventRecord(t->ev0, t->stream);
    if (e == cudaSuccess) {
      train_kernel<<<TRAIN_CTAS, TRAIN_THREADS, t->smem, t->stream>>>(t->K, t->KW, t->H, t->batch, t->n_rows, t->xbits_dev, t->y_dev,
                                                             t->theta_dev, t->state_dev, n_epochs, ord_dev, loss_dev, t->lr,
                                                             t->alpha, t->beta1, t->beta2, t->eps, t->tol, t->n_iter_no_change);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaEventRecord(t->ev1, t->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(loss_out, loss_dev, (size_t)n_epochs * 8, cudaMemcpyDeviceToHost, t->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(t->stream);
    if (e == cudaSuccess) e = cudaMemcpy(&after, t->state_dev, sizeof(after), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, t->ev0, t->ev1);
    cudaFree(ord_dev);
    cudaFree(loss_dev);
    if (e != cudaSuccess) { so_set_error("training launch failed: %s", cudaGetErrorString(e)); return SO_ECUDA; }
  }
  if (status) {
    status->epochs_done = after.n_iter - before.n_iter;
    status->n_iter = after.n_iter;
    status->stopped = after.stopped;
    status->adam_steps = after.t;
    status->best_loss = after.best;
    status->kernel_ms = ms;
  }
  return SO_OK;
}

int so_trainer_predict(so_trainer *t, int64_t n, const uint8_t *X, double *out) {
  SO_REQUIRE(t && (n == 0 || (X && out)), "NULL argument");
  SO_REQUIRE(n >= 0, "n=%lld is negative", (long long)n);
  if (n == 0) return SO_OK;
  SO_CUDA(cudaSetDevice(t->device));
  uint8_t *x_dev = nullptr;
  double *o_dev = nullptr;
  if (cudaMalloc(&x_dev, (size_t)n * t->K) != cudaSuccess || cudaMalloc(&o_dev, (size_t)n * 8) != cudaSuccess) {
    cudaFree(x_dev);
    cudaGetLastError();
    so_set_error("out of device memory for %lld rows", (long long)n);
    return SO_ENOMEM;
  }
  cudaError_t e = cudaMemcpyAsync(x_dev, X, (size_t)n * t->K, cudaMemcpyHostToDevice, t->stream);
  if (e == cudaSuccess) {
    predict_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, t->stream>>>(t->K, t->H, n, x_dev, t->theta_dev, o_dev);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, o_dev, (size_t)n * 8, cudaMemcpyDeviceToHost, t->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(t->stream);
  cudaFree(x_dev);
  cudaFree(o_dev);
  if (e != cudaSuccess) { so_set_error("predict failed: %s", cudaGetErrorString(e)); return SO_ECUDA; }
  return SO_OK;
}

int so_trainer_get(so_trainer *t, double *W1, double *b1, double *w2, double *b2) {
  SO_REQUIRE(t && W1 && b1 && w2 && b2, "NULL argument");
  SO_CUDA(cudaSetDevice(t->device));
  const size_t P = (size_t)(t->K + 1) * t->H + t->H + 1;
  std::vector<double> theta(P);
  SO_CUDA(cudaMemcpy(theta.data(), t->theta_dev, P * 8, cudaMemcpyDeviceToHost));
  memcpy(W1, theta.data(), (size_t)t->K * t->H * 8);
  memcpy(b1, theta.data() + (size_t)t->K * t->H, (size_t)t->H * 8);
  memcpy(w2, theta.data() + (size_t)(t->K + 1) * t->H, (size_t)t->H * 8);
  *b2 = theta[P - 1];
  return SO_OK;
}

}  // extern "C"
