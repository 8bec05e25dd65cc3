/*
 * ORACLE (test infrastructure, NOT product code) -- C restatement of the simulated-annealing loop in the exact
 * fixed-point arithmetic of the CUDA path, for parity checks at sizes the NumPy oracle (oracle/sa.py: anneal_fixed,
 * which this file mirrors statement by statement) cannot finish in seconds.  One chain per OpenMP thread.
 *
 * Reference followed: OML/optimize_SA.py:59-132 (proposal, incremental update, objective, Metropolis, commit) and
 * OML/train_surrogate.py:34-54 (gate, MLP, inverse scaling, sum).  Pinned against oracle/sa.py in
 * tests/test_oracle_c.py, which in turn is pinned against the literal restatement (optimize_faithful) and the
 * committed golden runs.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may load this.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SPLIT_BITS 20
#define SPLIT_MASK ((1LL << SPLIT_BITS) - 1)

typedef struct {
  int d, K, Hp;
  const int32_t *offsets; /* [K][2] */
  const int32_t *W1q;     /* [K][Hp] */
  const int32_t *b1q, *w2q;
  double c, b2, sigma, mu;
  int n_tree;
  const int32_t *feature, *left, *right;
  const double *thr;
  const uint8_t *active;
} model_t;

static inline int wrapi(int v, int d) { v %= d; return v < 0 ? v + d : v; }

/* pinned conversion (oracle/surrogate.py real_from_sums); compile with -ffp-contract=off */
static double real_from_sums(const model_t *m, int64_t hi, int64_t lo, int n) {
  double S = (double)hi * 1048576.0 + (double)lo;
  double nn = (double)n;
  double t3 = m->c * S + m->b2 * nn;
  return m->sigma * t3 + m->mu * nn;
}

static int gate_of(const model_t *m, const uint8_t *x, int r1, int r2, int flip) {
  if (m->n_tree == 0) return 1;
  int node = 0, d = m->d;
  while (m->feature[node] >= 0) {
    int f = m->feature[node];
    int v = wrapi(r2 + m->offsets[2 * f + 1], d) * d + wrapi(r1 + m->offsets[2 * f], d);
    int bit = x[v] ^ (v == flip);
    node = ((double)bit <= m->thr[node]) ? m->left[node] : m->right[node];
  }
  return m->active[node];
}

static int64_t row_sum(const model_t *m, const int32_t *h) {
  int64_t A = 0;
  for (int j = 0; j < m->Hp; ++j) A += (int64_t)m->w2q[j] * (int64_t)(h[j] > 0 ? h[j] : 0);
  return A;
}

static void build_state(const model_t *m, const uint8_t *x, int32_t *hq, int64_t *hi, int64_t *lo, int *nact) {
  int d = m->d, N = d * d;
  int64_t shi = 0, slo = 0;
  int n = 0;
  for (int r = 0; r < N; ++r) {
    int r1 = r % d, r2 = r / d;
    int32_t *h = hq + (size_t)r * m->Hp;
    for (int j = 0; j < m->Hp; ++j) h[j] = m->b1q[j];
    for (int k = 0; k < m->K; ++k) {
      int v = wrapi(r2 + m->offsets[2 * k + 1], d) * d + wrapi(r1 + m->offsets[2 * k], d);
      if (x[v])
        for (int j = 0; j < m->Hp; ++j) h[j] += m->W1q[(size_t)k * m->Hp + j];
    }
    if (gate_of(m, x, r1, r2, -1)) {
      int64_t A = row_sum(m, h);
      shi += A >> SPLIT_BITS;
      slo += A & SPLIT_MASK;
      ++n;
    }
  }
  shi += slo >> SPLIT_BITS;
  slo &= SPLIT_MASK;
  *hi = shi; *lo = slo; *nact = n;
}

/* Returns 0, or -1 on allocation failure.  x is updated in place; any output pointer may be NULL. */
int oracle_anneal_fixed(int d, int K, int Hp, const int32_t *offsets, const int32_t *W1q, const int32_t *b1q,
                        const int32_t *w2q, double c, double b2, double sigma, double mu, int n_tree,
                        const int32_t *feature, const double *thr, const int32_t *left, const int32_t *right,
                        const uint8_t *active, int64_t n_chains, int64_t n_steps, uint8_t *x, const double *temps,
                        const double *tscale, const int32_t *prop, const float *u, uint8_t *accept_out, int64_t *hi_out,
                        int64_t *lo_out, int32_t *nact_out, double *of_out) {
  model_t m = {d, K, Hp, offsets, W1q, b1q, w2q, c, b2, sigma, mu, n_tree, feature, left, right, thr, active};
  const int N = d * d;
  int failed = 0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t ch = 0; ch < n_chains; ++ch) {
    int32_t *hq = (int32_t *)malloc((size_t)N * Hp * sizeof(int32_t));
    int32_t *hnew = (int32_t *)malloc((size_t)K * Hp * sizeof(int32_t));
    int *rows = (int *)malloc((size_t)K * sizeof(int));
    if (!hq || !hnew || !rows) { failed = 1; free(hq); free(hnew); free(rows); continue; }
    uint8_t *xc = x + ch * N;
    int64_t hi, lo;
    int nact;
    build_state(&m, xc, hq, &hi, &lo, &nact);
    const double ts = tscale ? tscale[ch] : 1.0;
    for (int64_t s = 0; s < n_steps; ++s) {
      const double T = ts * temps[s];
      const int f = prop[ch * n_steps + s];
      const int f1 = f % d, f2 = f / d;
      const int sgn = xc[f] ? -1 : 1;
      int64_t dhi = 0, dlo = 0;
      int dn = 0;
      for (int k = 0; k < K; ++k) {
        int r1 = wrapi(f1 - offsets[2 * k], d), r2 = wrapi(f2 - offsets[2 * k + 1], d);
        int r = r2 * d + r1;
        rows[k] = r;
        const int32_t *h = hq + (size_t)r * Hp;
        int32_t *hn = hnew + (size_t)k * Hp;
        for (int j = 0; j < Hp; ++j) hn[j] = h[j] + sgn * W1q[(size_t)k * Hp + j];
        int g0 = gate_of(&m, xc, r1, r2, -1), g1 = gate_of(&m, xc, r1, r2, f);
        if (g1) { int64_t A = row_sum(&m, hn); dhi += A >> SPLIT_BITS; dlo += A & SPLIT_MASK; }
        if (g0) { int64_t A = row_sum(&m, h); dhi -= A >> SPLIT_BITS; dlo -= A & SPLIT_MASK; }
        dn += g1 - g0;
      }
      const double delta = real_from_sums(&m, dhi, dlo, dn);
      int acc;
      if (delta > 0.0) acc = 1;
      else if (T > 0.0) acc = exp(delta / T) > (double)u[ch * n_steps + s];
      else acc = 0;
      if (acc) {
        xc[f] ^= 1;
        for (int k = 0; k < K; ++k) memcpy(hq + (size_t)rows[k] * Hp, hnew + (size_t)k * Hp, (size_t)Hp * sizeof(int32_t));
        lo += dlo;
        hi += dhi + (lo >> SPLIT_BITS);
        lo &= SPLIT_MASK;
        nact += dn;
      }
      if (accept_out) accept_out[ch * n_steps + s] = (uint8_t)acc;
    }
    if (hi_out) hi_out[ch] = hi;
    if (lo_out) lo_out[ch] = lo;
    if (nact_out) nact_out[ch] = nact;
    if (of_out) of_out[ch] = real_from_sums(&m, hi, lo, nact);
    free(hq); free(hnew); free(rows);
  }
  return failed ? -1 : 0;
}

/* ---- the reference's fp64 rule with its ORIGINAL weights (no fixed point): decision oracle for full-size runs ----------
 * Same loop as above (OML/optimize_SA.py:59-132) with delta = sum over the K affected rows of
 * g'(y'*sigma + mu) - g(y*sigma + mu), y = w2 . relu(h) + b2 (OML/train_surrogate.py:40-54), h kept per chain in fp64 and
 * updated on accept.  delta equals the reference's OF_new - OF up to fp64 rounding (the reference subtracts two full
 * sums).  margin = |exp(delta/T) - u| of every decision taken in the Metropolis branch; per chain the minimum and the
 * number of decisions with margin < margin_thr are returned so that callers can state how close any tie was. */
static int gate_of_fp(int n_tree, const int32_t *feature, const double *thr, const int32_t *left, const int32_t *right,
                      const uint8_t *active, const int32_t *offsets, const uint8_t *x, int d, int r1, int r2, int flip) {
  if (n_tree == 0) return 1;
  int node = 0;
  while (feature[node] >= 0) {
    int f = feature[node];
    int v = wrapi(r2 + offsets[2 * f + 1], d) * d + wrapi(r1 + offsets[2 * f], d);
    int bit = x[v] ^ (v == flip);
    node = ((double)bit <= thr[node]) ? left[node] : right[node];
  }
  return active[node];
}

int oracle_anneal_fp64(int d, int K, int H, const int32_t *offsets, const double *W1, const double *b1, const double *w2,
                       double b2, double sigma, double mu, int n_tree, const int32_t *feature, const double *thr,
                       const int32_t *left, const int32_t *right, const uint8_t *active, int64_t n_chains,
                       int64_t n_steps, uint8_t *x, const double *temps, const double *tscale, const int32_t *prop,
                       const float *u, double margin_thr, uint8_t *accept_out, double *min_margin_out,
                       int64_t *n_below_out, double *of_out) {
  const int N = d * d;
  int failed = 0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t ch = 0; ch < n_chains; ++ch) {
    double *h = (double *)malloc((size_t)N * H * sizeof(double));
    if (!h) { failed = 1; continue; }
    uint8_t *xc = x + ch * N;
    for (int r = 0; r < N; ++r) {
      int r1 = r % d, r2 = r / d;
      double *hr = h + (size_t)r * H;
      for (int j = 0; j < H; ++j) hr[j] = b1[j];
      for (int k = 0; k < K; ++k) {
        int v = wrapi(r2 + offsets[2 * k + 1], d) * d + wrapi(r1 + offsets[2 * k], d);
        if (xc[v])
          for (int j = 0; j < H; ++j) hr[j] += W1[(size_t)k * H + j];
      }
    }
    const double ts = tscale ? tscale[ch] : 1.0;
    double min_margin = INFINITY;
    int64_t n_below = 0;
    for (int64_t s = 0; s < n_steps; ++s) {
      const double T = ts * temps[s];
      const int f = prop[ch * n_steps + s];
      const int f1 = f % d, f2 = f / d;
      const double sgn = xc[f] ? -1.0 : 1.0;
      double s_old = 0.0, s_new = 0.0;
      for (int k = 0; k < K; ++k) {
        int r1 = wrapi(f1 - offsets[2 * k], d), r2 = wrapi(f2 - offsets[2 * k + 1], d);
        const double *hr = h + (size_t)(r2 * d + r1) * H;
        int g0 = gate_of_fp(n_tree, feature, thr, left, right, active, offsets, xc, d, r1, r2, -1);
        int g1 = gate_of_fp(n_tree, feature, thr, left, right, active, offsets, xc, d, r1, r2, f);
        if (!g0 && !g1) continue;
        double y0 = b2, y1 = b2;
        for (int j = 0; j < H; ++j) {
          double a = hr[j], b = a + sgn * W1[(size_t)k * H + j];
          y0 += w2[j] * (a > 0.0 ? a : 0.0);
          y1 += w2[j] * (b > 0.0 ? b : 0.0);
        }
        if (g0) s_old += y0 * sigma + mu;
        if (g1) s_new += y1 * sigma + mu;
      }
      const double delta = s_new - s_old;
      int acc;
      if (delta > 0.0) acc = 1;
      else if (T > 0.0) {
        const double pr = exp(delta / T), uu = (double)u[ch * n_steps + s];
        acc = pr > uu;
        const double mg = fabs(pr - uu);
        if (mg < min_margin) min_margin = mg;
        if (mg < margin_thr) ++n_below;
      } else acc = 0;
      if (acc) {
        xc[f] ^= 1;
        for (int k = 0; k < K; ++k) {
          int r1 = wrapi(f1 - offsets[2 * k], d), r2 = wrapi(f2 - offsets[2 * k + 1], d);
          double *hr = h + (size_t)(r2 * d + r1) * H;
          for (int j = 0; j < H; ++j) hr[j] += sgn * W1[(size_t)k * H + j];
        }
      }
      if (accept_out) accept_out[ch * n_steps + s] = (uint8_t)acc;
    }
    if (min_margin_out) min_margin_out[ch] = min_margin;
    if (n_below_out) n_below_out[ch] = n_below;
    if (of_out) {                                   /* final objective from a fresh fp64 evaluation */
      double of = 0.0;
      for (int r = 0; r < N; ++r) {
        if (!gate_of_fp(n_tree, feature, thr, left, right, active, offsets, xc, d, r % d, r / d, -1)) continue;
        const double *hr = h + (size_t)r * H;
        double y = b2;
        for (int j = 0; j < H; ++j) y += w2[j] * (hr[j] > 0.0 ? hr[j] : 0.0);
        of += y * sigma + mu;
      }
      of_out[ch] = of;
    }
    free(h);
  }
  return failed ? -1 : 0;
}

/* LSC / NH3-exported site types of one structure by the closed-form local rules (oracle/site_types.py) are NOT
 * duplicated here: the NumPy oracle is fast enough for them at every tested size. */
