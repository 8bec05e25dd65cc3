// so_tree.cu -- the classifier side of the surrogate's training on the device:
//
//   so_tree_fit             DecisionTreeClassifier(max_depth=30, random_state=0).fit(X_train, y_train)
//                                                                              OML/train_surrogate.py:111-114
//   so_build_training_set   all_syms = generate_all_translations_and_rotations() of every structure, the site rates
//                           tiled x3 and the 0.06 * max activity threshold       OML/dynamic_cat.py:270-322,
//                                                                              OML/train_surrogate.py:72-98
//
// The tree is scikit-learn's CART for 0/1 features (sklearn/tree/_tree.pyx DepthFirstTreeBuilder, _splitter.pyx
// node_split_best, _criterion.pyx Gini, utils/_random.pxd our_rand_r; restated in oracle/tree.py and pinned there
// against the installed scikit-learn): the same depth-first build order, the same Fisher-Yates feature draw with the
// constant-feature bookkeeping (it decides ties between equally good splits), the same float64 expressions -- so the
// fitted tree is node for node the one scikit-learn returns.
//
// Mapping: the training matrix is stored as one bit column per feature (bit s of column f = X[s][f]); a node is a bit
// mask over the samples.  One persistent CTA builds the whole tree: per node all 32 warps count, for every feature,
// the samples of the node with the feature set and those of class 1 among them (popc of col & mask [& y]); thread 0
// replays scikit-learn's sequential feature loop on these counts; all threads then split the mask for the two children
// (right child pushed first, left child built first).  No host round trip per node.
#include "so_internal.cuh"
#include <math.h>
#include <vector>

namespace {

constexpr int TREE_THREADS = 1024;
constexpr int TREE_STACK = 64;            // DFS stack entries (depth <= 62)

struct StackRec {
  int depth, parent, is_left, n_known;
  double impurity;
};

__device__ __forceinline__ uint32_t xorshift_rand_r(uint32_t &s) {
  if (s == 0) s = 1;                      // DEFAULT_SEED
  s ^= s << 13;
  s ^= s >> 17;
  s ^= s << 5;
  return s % 2147483648u;                 // RAND_R_MAX + 1
}

// gini of the two children in scikit-learn's operation order, no contraction
__device__ __forceinline__ void gini_children(double l0, double l1, double r0, double r1, double &gl, double &gr) {
  const double wl = __dadd_rn(l0, l1), wr = __dadd_rn(r0, r1);
  const double sql = __dadd_rn(__dadd_rn(0.0, __dmul_rn(l0, l0)), __dmul_rn(l1, l1));
  const double sqr = __dadd_rn(__dadd_rn(0.0, __dmul_rn(r0, r0)), __dmul_rn(r1, r1));
  gl = __dsub_rn(1.0, __ddiv_rn(sql, __dmul_rn(wl, wl)));
  gr = __dsub_rn(1.0, __ddiv_rn(sqr, __dmul_rn(wr, wr)));
}

// pack X[n][K] bytes into bit columns col[K][Wn]
__global__ void pack_columns_kernel(const uint8_t *__restrict__ X, long long n, int K, int Wn, uint32_t *__restrict__ col) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)K * Wn) return;
  const int f = (int)(t / Wn), w = (int)(t - (long long)f * Wn);
  uint32_t bits = 0;
  for (int b = 0; b < 32; ++b) {
    long long s = (long long)w * 32 + b;
    if (s < n && X[s * K + f]) bits |= 1u << b;
  }
  col[t] = bits;
}

__global__ void __launch_bounds__(TREE_THREADS, 1)
    tree_fit_kernel(const uint32_t *__restrict__ col, const uint32_t *__restrict__ ybits, long long n, int K, int Wn,
                    int two_classes, int max_depth, uint32_t seed, int max_nodes, uint32_t *__restrict__ masks,
                    int32_t *feature, double *thr, int32_t *left, int32_t *right, double *count0, double *count1,
                    int32_t *status /* [0] = node count, [1] = error flag */) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  int *s_ones = (int *)s_raw;               // [K] samples of the node with the feature set
  int *s_ones1 = s_ones + K;                // [K] ... of class 1
  int *s_features = s_ones1 + K;            // [K] scikit-learn's `features` permutation (persists across nodes)
  int *s_const = s_features + K;            // [K] `constant_features`
  __shared__ StackRec s_stack[TREE_STACK];
  __shared__ int s_top, s_nodes, s_best, s_leaf, s_n_node, s_n1, s_ntotal;
  __shared__ double s_gl, s_gr;
  __shared__ uint32_t s_rng;
  __shared__ int s_red[2][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;

  for (int f = tid; f < K; f += blockDim.x) { s_features[f] = f; s_const[f] = 0; }
  // root mask: all samples
  for (int w = tid; w < Wn; w += blockDim.x) {
    const long long rest = n - (long long)w * 32;
    masks[w] = rest >= 32 ? 0xffffffffu : (rest <= 0 ? 0u : ((1u << rest) - 1u));
  }
  if (tid == 0) {
    s_stack[0] = StackRec{0, -1, 0, 0, INFINITY};
    s_top = 1; s_nodes = 0; s_rng = seed;
  }
  __syncthreads();
  bool first = true;
  while (s_top > 0) {
    const int top = s_top - 1;                          // the popped record keeps its slot until the children are pushed
    const StackRec rec = s_stack[top];
    const uint32_t *mask = masks + (size_t)top * Wn;
    // ---- node size and class-1 count ---------------------------------------------------------------------
    int cn = 0, c1 = 0;
    for (int w = tid; w < Wn; w += blockDim.x) {
      const uint32_t m = mask[w];
      cn += __popc(m);
      c1 += __popc(m & ybits[w]);
    }
    cn = __reduce_add_sync(SO_FULL, cn);
    c1 = __reduce_add_sync(SO_FULL, c1);
    if (lane == 0) { s_red[0][warp] = cn; s_red[1][warp] = c1; }
    __syncthreads();
    if (tid == 0) {
      int a = 0, b = 0;
      for (int w = 0; w < nwarp; ++w) { a += s_red[0][w]; b += s_red[1][w]; }
      s_n_node = a; s_n1 = two_classes ? b : 0;
    }
    __syncthreads();
    const int n_node = s_n_node;
    const double n1 = (double)s_n1, n0 = (double)(n_node - s_n1);
    double impurity = rec.impurity;
    if (first) {                                        // root impurity (DepthFirstTreeBuilder: `if first`)
      const double sq = __dadd_rn(__dadd_rn(0.0, __dmul_rn(n0, n0)), __dmul_rn(n1, n1));
      impurity = __dsub_rn(1.0, __ddiv_rn(sq, __dmul_rn((double)n_node, (double)n_node)));
      first = false;
    }
    bool is_leaf = rec.depth >= max_depth || n_node < 2 || impurity <= 2.220446049250313e-16;
    if (!is_leaf) {
      // ---- per-feature counts: one warp per feature ---------------------------------------------------------
      for (int f = warp; f < K; f += nwarp) {
        const uint32_t *cf = col + (size_t)f * Wn;
        int a = 0, b = 0;
        for (int w = lane; w < Wn; w += 32) {
          const uint32_t m = mask[w] & cf[w];
          a += __popc(m);
          b += __popc(m & ybits[w]);
        }
        a = __reduce_add_sync(SO_FULL, a);
        b = __reduce_add_sync(SO_FULL, b);
        if (lane == 0) { s_ones[f] = a; s_ones1[f] = two_classes ? b : 0; }
      }
      __syncthreads();
      // ---- node_split_best, sequential (the feature order and the strict '>' decide ties) ---------------------
      if (tid == 0) {
        const int n_known = rec.n_known;
        int n_total = n_known, f_i = K, n_found = 0, n_drawn = 0, n_visited = 0, best = -1;
        double best_proxy = -INFINITY, bgl = 0.0, bgr = 0.0;
        uint32_t rng = s_rng;
        while (f_i > n_total && (n_visited < K || n_visited <= n_found + n_drawn)) {
          ++n_visited;
          int f_j = n_drawn + (int)(xorshift_rand_r(rng) % (uint32_t)(f_i - n_found - n_drawn));
          if (f_j < n_known) {
            const int t = s_features[n_drawn]; s_features[n_drawn] = s_features[f_j]; s_features[f_j] = t;
            ++n_drawn;
            continue;
          }
          f_j += n_found;
          const int f = s_features[f_j];
          const int on = s_ones[f];
          if (on == 0 || on == n_node) {                // constant in this node
            const int t = s_features[f_j]; s_features[f_j] = s_features[n_total]; s_features[n_total] = t;
            ++n_found; ++n_total;
            continue;
          }
          --f_i;
          { const int t = s_features[f_i]; s_features[f_i] = s_features[f_j]; s_features[f_j] = t; }
          const double r1 = (double)s_ones1[f], r0 = (double)(on - s_ones1[f]);
          const double l1 = __dsub_rn(n1, r1), l0 = __dsub_rn(n0, r0);
          double gl, gr;
          gini_children(l0, l1, r0, r1, gl, gr);
          const double proxy = __dsub_rn(__dmul_rn(-__dadd_rn(r0, r1), gr), __dmul_rn(__dadd_rn(l0, l1), gl));
          if (proxy > best_proxy) { best_proxy = proxy; best = f; bgl = gl; bgr = gr; }
        }
        for (int i = 0; i < n_known; ++i) s_features[i] = s_const[i];
        for (int i = 0; i < n_found; ++i) s_const[n_known + i] = s_features[n_known + i];
        s_rng = rng; s_best = best; s_gl = bgl; s_gr = bgr; s_ntotal = n_total;
      }
      __syncthreads();
      is_leaf = s_best < 0;
    }
    // ---- add the node (ids in build order) -----------------------------------------------------------------------
    const int node_id = s_nodes;
    if (node_id >= max_nodes) {
      if (tid == 0) status[1] = 1;
      break;
    }
    if (tid == 0) {
      if (rec.parent >= 0) (rec.is_left ? left : right)[rec.parent] = node_id;
      feature[node_id] = is_leaf ? -2 : s_best;
      thr[node_id] = is_leaf ? -2.0 : 0.5;
      left[node_id] = -1; right[node_id] = -1;
      count0[node_id] = n0; count1[node_id] = n1;
    }
    __syncthreads();
    if (is_leaf) {
      if (tid == 0) { s_top = top; s_nodes = node_id + 1; }
    } else {
      if (top + 2 > TREE_STACK) {
        if (tid == 0) status[1] = 2;
        break;
      }
      // children masks: right child takes the popped slot, left child the slot above it (popped next)
      const uint32_t *cf = col + (size_t)s_best * Wn;
      uint32_t *mr = masks + (size_t)top * Wn, *ml = masks + (size_t)(top + 1) * Wn;
      for (int w = tid; w < Wn; w += blockDim.x) {
        const uint32_t m = mr[w], c = cf[w];
        ml[w] = m & ~c;
        mr[w] = m & c;
      }
      if (tid == 0) {
        s_stack[top] = StackRec{rec.depth + 1, node_id, 0, s_ntotal, s_gr};
        s_stack[top + 1] = StackRec{rec.depth + 1, node_id, 1, s_ntotal, s_gl};
        s_top = top + 2; s_nodes = node_id + 1;
      }
    }
    __syncthreads();
  }
  if (tid == 0) status[0] = s_nodes;
}

// ---- training set: all translations x 3 rotations of every structure + activity labels ---------------------------
// row (structure s, site r, rotation k) = 3*(s*N + r) + k; column v = x_s[translate by r, then rotate by k] as
// OML/dynamic_cat.py:270-322: translated[v] = x[idx(v1 + r1, v2 + r2)]; k = 1: new[(d1, d2)] = old[(-d1 - d2, d1)];
// k = 2: new[(d1, d2)] = old[(d2, -d1 - d2)].
__global__ void symmetry_rows_kernel(const uint32_t *__restrict__ bits, long long n_struct, int d, int N, int W,
                                     uint8_t *__restrict__ X) {
  const long long row = blockIdx.x;
  const long long s = row / (3LL * N);
  const int rem = (int)(row - s * 3LL * N), r = rem / 3, k = rem - 3 * r;
  const int r1 = r % d, r2 = r / d;
  const uint32_t *w = bits + s * W;
  for (int v = threadIdx.x; v < N; v += blockDim.x) {
    const int d1 = v % d, d2 = v / d;
    int a1, a2;                                    // position read in the TRANSLATED vector
    if (k == 0) { a1 = d1; a2 = d2; }
    else if (k == 1) { a1 = -d1 - d2; a2 = d1; }
    else { a1 = d2; a2 = -d1 - d2; }
    a1 %= d; if (a1 < 0) a1 += d;
    a2 %= d; if (a2 < 0) a2 += d;
    const int src = ((a2 + r2) % d) * d + (a1 + r1) % d;
    X[row * N + v] = (w[src >> 5] >> (src & 31)) & 1u;
  }
}

__global__ void rate_max_kernel(const double *__restrict__ rates, long long n, double *out) {
  __shared__ double s[256];
  double m = -INFINITY;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, rates[i]);
  s[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] = fmax(s[threadIdx.x], s[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = s[0];
}

// site_is_active of the x3-tiled rates (train_surrogate.py:78-95): row 3*i + k is active iff rates[i] > frac * max
__global__ void activity_kernel(const double *__restrict__ rates, long long n, double frac, const double *y_max,
                                uint8_t *__restrict__ active) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= 3 * n) return;
  active[t] = rates[t / 3] > __dmul_rn(frac, *y_max) ? 1 : 0;
}

}  // namespace

int so_tree_fit(int device, int64_t n_rows, int K, const uint8_t *X, const uint8_t *y, int two_classes, int max_depth,
                uint32_t splitter_seed, int max_nodes, int32_t *feature, double *thr, int32_t *left, int32_t *right,
                double *count0, double *count1, int32_t *n_nodes, float *kernel_ms) {
  SO_REQUIRE(X && y && feature && thr && left && right && count0 && count1 && n_nodes, "NULL argument");
  SO_REQUIRE(n_rows >= 1 && n_rows < (1LL << 31) && K >= 1 && K <= 4096, "training set shape %lld x %d not supported",
             (long long)n_rows, K);
  SO_REQUIRE(max_depth >= 0 && max_depth <= TREE_STACK - 2, "max_depth %d outside [0, %d]", max_depth, TREE_STACK - 2);
  SO_REQUIRE(max_nodes >= 1, "max_nodes must be positive");
  SO_CUDA(cudaSetDevice(device));
  const int Wn = (int)((n_rows + 31) / 32);
  const size_t smem = (size_t)4 * K * sizeof(int);
  DevBuf dX, dy, dcol, dyb, dmask, dfe, dth, dl, dr, dc0, dc1, dst;
  SO_TRY(dX.alloc((size_t)n_rows * K));
  SO_TRY(dy.alloc((size_t)n_rows));
  SO_TRY(dcol.alloc((size_t)K * Wn * 4));
  SO_TRY(dyb.alloc((size_t)Wn * 4));
  SO_TRY(dmask.alloc((size_t)TREE_STACK * Wn * 4));
  SO_TRY(dfe.alloc((size_t)max_nodes * 4));
  SO_TRY(dth.alloc((size_t)max_nodes * 8));
  SO_TRY(dl.alloc((size_t)max_nodes * 4));
  SO_TRY(dr.alloc((size_t)max_nodes * 4));
  SO_TRY(dc0.alloc((size_t)max_nodes * 8));
  SO_TRY(dc1.alloc((size_t)max_nodes * 8));
  SO_TRY(dst.alloc(16));
  cudaStream_t st = 0;
  SO_CUDA(cudaMemcpyAsync(dX.p, X, (size_t)n_rows * K, cudaMemcpyHostToDevice, st));
  SO_CUDA(cudaMemcpyAsync(dy.p, y, (size_t)n_rows, cudaMemcpyHostToDevice, st));
  SO_CUDA(cudaMemsetAsync(dst.p, 0, 16, st));
  pack_columns_kernel<<<(unsigned)(((long long)K * Wn + 255) / 256), 256, 0, st>>>(dX.as<uint8_t>(), n_rows, K, Wn,
                                                                                   dcol.as<uint32_t>());
  pack_columns_kernel<<<(unsigned)((Wn + 255) / 256), 256, 0, st>>>(dy.as<uint8_t>(), n_rows, 1, Wn, dyb.as<uint32_t>());
  SO_CUDA(cudaGetLastError());
  cudaEvent_t e0, e1;
  SO_CUDA(cudaEventCreate(&e0));
  SO_CUDA(cudaEventCreate(&e1));
  SO_CUDA(cudaFuncSetAttribute(tree_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SO_CUDA(cudaEventRecord(e0, st));
  tree_fit_kernel<<<1, TREE_THREADS, smem, st>>>(dcol.as<uint32_t>(), dyb.as<uint32_t>(), n_rows, K, Wn, two_classes,
                                                 max_depth, splitter_seed, max_nodes, dmask.as<uint32_t>(),
                                                 dfe.as<int32_t>(), dth.as<double>(), dl.as<int32_t>(), dr.as<int32_t>(),
                                                 dc0.as<double>(), dc1.as<double>(), dst.as<int32_t>());
  SO_CUDA(cudaGetLastError());
  SO_CUDA(cudaEventRecord(e1, st));
  int32_t status[4] = {0, 0, 0, 0};
  SO_CUDA(cudaMemcpyAsync(status, dst.p, 16, cudaMemcpyDeviceToHost, st));
  SO_CUDA(cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (kernel_ms) *kernel_ms = ms;
  SO_REQUIRE(status[1] == 0, status[1] == 1 ? "the tree needs more than max_nodes = %d nodes" : "tree deeper than the stack (%d)",
             status[1] == 1 ? max_nodes : TREE_STACK);
  const int nn = status[0];
  *n_nodes = nn;
  SO_CUDA(cudaMemcpy(feature, dfe.p, (size_t)nn * 4, cudaMemcpyDeviceToHost));
  SO_CUDA(cudaMemcpy(thr, dth.p, (size_t)nn * 8, cudaMemcpyDeviceToHost));
  SO_CUDA(cudaMemcpy(left, dl.p, (size_t)nn * 4, cudaMemcpyDeviceToHost));
  SO_CUDA(cudaMemcpy(right, dr.p, (size_t)nn * 4, cudaMemcpyDeviceToHost));
  SO_CUDA(cudaMemcpy(count0, dc0.p, (size_t)nn * 8, cudaMemcpyDeviceToHost));
  SO_CUDA(cudaMemcpy(count1, dc1.p, (size_t)nn * 8, cudaMemcpyDeviceToHost));
  return SO_OK;
}

int so_build_training_set(so_lattice *L, int64_t n_struct, const uint32_t *bits, const double *site_rates, double frac,
                          uint8_t *X, uint8_t *active, double *y_max) {
  SO_REQUIRE(L && bits, "NULL argument");
  SO_REQUIRE(n_struct >= 1, "no structures");
  SO_REQUIRE(!active || site_rates, "activity labels need the site rates");
  SO_CUDA(cudaSetDevice(L->device));
  const int N = L->N;
  const long long rows = n_struct * 3LL * N;
  DevBuf db, dx, dr, dm, da;
  SO_TRY(db.alloc((size_t)n_struct * L->W * 4));
  SO_CUDA(cudaMemcpy(db.p, bits, (size_t)n_struct * L->W * 4, cudaMemcpyHostToDevice));
  if (X) {
    const long long slab_rows = std::max<long long>(3LL * N, ((512LL << 20) / N) / (3LL * N) * (3LL * N));
    SO_TRY(dx.alloc((size_t)std::min(slab_rows, rows) * N));
    for (long long r0 = 0; r0 < rows; r0 += slab_rows) {
      const long long nr = std::min(slab_rows, rows - r0);
      const long long s0 = r0 / (3LL * N);
      symmetry_rows_kernel<<<(unsigned)nr, 128>>>(db.as<uint32_t>() + s0 * L->W, nr / (3LL * N), L->d, N, L->W,
                                                  dx.as<uint8_t>());
      SO_CUDA(cudaGetLastError());
      SO_CUDA(cudaMemcpy(X + r0 * N, dx.p, (size_t)nr * N, cudaMemcpyDeviceToHost));
    }
  }
  if (site_rates) {
    const long long n = n_struct * (long long)N;
    SO_TRY(dr.alloc((size_t)n * 8));
    SO_TRY(dm.alloc(8));
    SO_CUDA(cudaMemcpy(dr.p, site_rates, (size_t)n * 8, cudaMemcpyHostToDevice));
    rate_max_kernel<<<1, 256>>>(dr.as<double>(), n, dm.as<double>());
    if (active) {
      SO_TRY(da.alloc((size_t)3 * n));
      activity_kernel<<<(unsigned)((3 * n + 255) / 256), 256>>>(dr.as<double>(), n, frac, dm.as<double>(), da.as<uint8_t>());
      SO_CUDA(cudaGetLastError());
      SO_CUDA(cudaMemcpy(active, da.p, (size_t)3 * n, cudaMemcpyDeviceToHost));
    }
    if (y_max) SO_CUDA(cudaMemcpy(y_max, dm.p, 8, cudaMemcpyDeviceToHost));
  }
  SO_CUDA(cudaDeviceSynchronize());
  return SO_OK;
}
