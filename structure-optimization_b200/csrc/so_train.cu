// so_train.cu -- the surrogate's regressor trained on the device: the K -> H -> 1 ReLU MLP of
// OML/train_surrogate.py:141-152, mini-batch Adam with L2 penalty exactly as scikit-learn's MLPRegressor runs it
// (squared loss / 2, alpha / 2 |coefs|^2 / batch, one Adam step per batch, tol-based stop on the epoch loss), in fp64.
//
// Training is a strictly sequential chain of tiny batches (200 rows x K x H: ~1 M fp64 FMA per step, up to ~1e6
// steps), so the whole optimiser state lives in the shared memory of ONE persistent CTA: weights, both Adam moments,
// the batch's activations and its rows as bit-vectors (inputs are 0/1 occupancies: the first layer is a sum of weight
// rows over set bits, its gradient a masked sum over the batch).  Nothing but the batch rows (K/8 B each) and the
// targets is read from HBM per step.  The order of the samples in every epoch is supplied by the caller, so a run
// can replay scikit-learn's own shuffles (parity tests) or any other stream.
#include "so_internal.cuh"
#include <math.h>
#include <string.h>

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

__device__ __forceinline__ double block_sum(double v, double *scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SO_FULL, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < (int)(blockDim.x >> 5) ? scratch[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SO_FULL, v, o);
    if (lane == 0) scratch[32] = v;
  }
  __syncthreads();
  return scratch[32];
}

// parameter vector: [ (K+1)*H : W1 rows 0..K-1, then b1 as row K ][ H : w2 ][ 1 : b2 ]
__global__ void __launch_bounds__(TRAIN_THREADS, 1)
    train_kernel(int K, int KW, int H, int batch, long long n_rows, const uint32_t *__restrict__ xbits,
                 const double *__restrict__ y, double *theta_g, TrainState *state_g, int n_epochs,
                 const int32_t *__restrict__ orders, double *loss_out, double lr, double alpha, double beta1, double beta2,
                 double eps, double tol, int n_iter_no_change) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int P1 = (K + 1) * H, P = P1 + H + 1;
  double *th = (double *)s_raw;            // parameters
  double *mo = th + P;                     // first moment
  double *ve = mo + P;                     // second moment
  double *act = ve + P;                    // [batch][H] activations, then delta1
  double *d2 = act + (size_t)batch * H;    // [batch]
  double *gw2 = d2 + batch;                // [H + 1] gradients of w2, b2
  double *scratch = gw2 + H + 1;           // [34]: block reduction + lr_t
  uint32_t *xb = (uint32_t *)(scratch + 34);   // [batch][KW]
  __shared__ TrainState st;
  __shared__ int s_stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < 3 * P; i += blockDim.x) th[i] = theta_g[i];
  if (tid == 0) { st = *state_g; s_stop = st.stopped; }
  __syncthreads();

  for (int e = 0; e < n_epochs && !s_stop; ++e) {
    const int32_t *order = orders + (size_t)e * n_rows;
    double acc_loss = 0.0;                 // thread 0 only
    for (long long s0 = 0; s0 < n_rows; s0 += batch) {
      const int nb = (int)min((long long)batch, n_rows - s0);
      const double inv_nb = 1.0 / (double)nb;
      // ---- batch rows (bit-vectors) and targets ----
      for (int i = tid; i < nb * KW; i += blockDim.x) {
        const int r = i / KW, w = i - r * KW;
        xb[i] = xbits[(size_t)order[s0 + r] * KW + w];
      }
      if (tid < nb) d2[tid] = y[order[s0 + tid]];
      if (tid == 0) {
        st.t += 1;
        scratch[33] = lr * sqrt(1.0 - pow(beta2, (double)st.t)) / (1.0 - pow(beta1, (double)st.t));
      }
      __syncthreads();
      // ---- forward, first layer: a[r][j] = relu(b1[j] + sum over set bits k of W1[k][j]) ----
      for (int p = tid; p < nb * H; p += blockDim.x) {
        const int r = p / H, j = p - r * H;
        double h = th[K * H + j];
        for (int w = 0; w < KW; ++w) {
          uint32_t bits = xb[r * KW + w];
          const double *row = th + (size_t)(32 * w) * H + j;
          while (bits) {
            const int k = __ffs(bits) - 1;
            bits &= bits - 1;
            h += row[k * H];
          }
        }
        act[p] = fmax(h, 0.0);
      }
      __syncthreads();
      // ---- output, residual, loss terms ----
      double part = 0.0;
      if (tid < nb) {
        double pred = th[P1 + H];
        for (int j = 0; j < H; ++j) pred += act[tid * H + j] * th[P1 + j];
        const double dd = pred - d2[tid];
        d2[tid] = dd;
        part = dd * dd;
      }
      const double sq = block_sum(part, scratch);
      part = 0.0;
      for (int i = tid; i < K * H; i += blockDim.x) part += th[i] * th[i];
      if (tid < H) part += th[P1 + tid] * th[P1 + tid];
      const double wsq = block_sum(part, scratch);
      if (tid == 0) acc_loss += (0.5 * sq * inv_nb + 0.5 * alpha * wsq * inv_nb) * nb;
      // ---- gradients of the output layer (warp j: column j of a^T delta2; warp H: sum of delta2) ----
      if (warp <= H) {
        double g = 0.0;
        for (int r = lane; r < nb; r += 32) g += (warp < H ? act[r * H + warp] : 1.0) * d2[r];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(SO_FULL, g, o);
        if (lane == 0) gw2[warp] = warp < H ? (g + alpha * th[P1 + warp]) / (double)nb : g / (double)nb;
      }
      __syncthreads();
      // ---- delta1 in place of the activations (old w2) ----
      for (int p = tid; p < nb * H; p += blockDim.x) {
        const int r = p / H, j = p - r * H;
        act[p] = act[p] > 0.0 ? d2[r] * th[P1 + j] : 0.0;
      }
      __syncthreads();
      // ---- first-layer gradients + Adam, one parameter per thread slot; then the output layer's parameters ----
      const double lr_t = scratch[33];
      for (int p = tid; p < P; p += blockDim.x) {
        double g;
        if (p < P1) {
          const int k = p / H, j = p - k * H;
          g = 0.0;
          if (k < K) {
            const int w = k >> 5, sh = k & 31;
            for (int r = 0; r < nb; ++r)
              if ((xb[r * KW + w] >> sh) & 1u) g += act[r * H + j];
            g = (g + alpha * th[p]) / (double)nb;
          } else {                                   // b1: no penalty
            for (int r = 0; r < nb; ++r) g += act[r * H + j];
            g /= (double)nb;
          }
        } else {
          g = gw2[p - P1];
        }
        const double m = beta1 * mo[p] + (1.0 - beta1) * g;
        const double v = beta2 * ve[p] + (1.0 - beta2) * (g * g);
        mo[p] = m;
        ve[p] = v;
        th[p] += -lr_t * m / (sqrt(v) + eps);
      }
      __syncthreads();
    }
    if (tid == 0) {
      const double cur = acc_loss / (double)n_rows;
      loss_out[e] = cur;
      st.n_iter += 1;
      st.count = cur > st.best - tol ? st.count + 1 : 0;
      if (cur < st.best) st.best = cur;
      if (st.count > n_iter_no_change) { st.stopped = 1; s_stop = 1; }
    }
    __syncthreads();
  }
  for (int i = tid; i < 3 * P; i += blockDim.x) theta_g[i] = th[i];
  if (tid == 0) *state_g = st;
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

size_t train_smem_bytes(int K, int H, int batch) {
  const size_t P = (size_t)(K + 1) * H + H + 1;
  const int KW = (K + 31) / 32;
  return (3 * P + (size_t)batch * H + batch + H + 1 + 34) * 8 + (size_t)batch * KW * 4 + 16;
}

}  // namespace

extern "C" {

int so_trainer_create(int device, int64_t n_rows, int K, int H, const uint8_t *X, const double *y, const double *W1,
                      const double *b1, const double *w2, double b2, const so_train_params *hp, so_trainer **out) {
  SO_REQUIRE(out, "out is NULL");
  *out = nullptr;
  SO_REQUIRE(X && y && W1 && b1 && w2 && hp, "NULL argument");
  SO_REQUIRE(n_rows >= 1 && n_rows < (1LL << 31), "n_rows=%lld out of range", (long long)n_rows);
  SO_REQUIRE(K >= 1 && H >= 1 && H <= 31, "K=%d, H=%d: need K >= 1 and 1 <= H <= 31", K, H);
  SO_REQUIRE(hp->batch_size >= 1 && hp->batch_size <= TRAIN_THREADS, "batch_size=%d out of range [1,%d]", hp->batch_size,
             TRAIN_THREADS);
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
      train_kernel<<<1, TRAIN_THREADS, t->smem, t->stream>>>(t->K, t->KW, t->H, t->batch, t->n_rows, t->xbits_dev, t->y_dev,
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
