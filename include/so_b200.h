/*
 * so_b200.h -- C ABI of the B200-native structure-optimisation hot path.
 *
 * The reference (VlachosGroup/Structure-Optimization) is pure Python and has no FFI layer; its boundary for
 * this path is the Python object API of the catalyst classes, `optimize()` and the `surrogate` class.  The
 * Python package in structure-optimization_b200/so_b200 mirrors that API and binds ONLY these entry points
 * (ctypes).  Each function below cites the reference routine it replaces (paths relative to the upstream
 * repository).  Conventions: every function returns 0 on success or a negative SO_E* code, and
 * so_last_error() returns a thread-local message.  All pointers are HOST pointers unless the name ends in
 * `_dev`.  Handles are opaque; one host thread per handle.  No borrowed pointer survives a call.
 */
#ifndef SO_B200_H
#define SO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SO_OK 0
#define SO_EINVAL (-1)  /* bad argument                                           */
#define SO_ECUDA  (-2)  /* CUDA runtime error (message in so_last_error)           */
#define SO_ENOMEM (-3)  /* host or device allocation failed                        */
#define SO_ESTATE (-4)  /* call sequence error (e.g. anneal before surrogate)      */

#define SO_KIND_LSC 0   /* OML/LSC_cat.py  */
#define SO_KIND_NH3 1   /* OML/NH3_cat.py  */

typedef struct so_lattice so_lattice;     /* slab topology tables on one device         */
typedef struct so_surrogate so_surrogate; /* quantised surrogate tables on one device   */
typedef struct so_chains so_chains;       /* state of n independent SA chains           */

int so_version(void);
const char *so_last_error(void);
int so_device_count(int *n_out);

/* ---- lattice: replaces the slab + graph construction of dynamic_cat.__init__ (OML/dynamic_cat.py:15-59),
 *      LSC_cat.__init__ (OML/LSC_cat.py:19-82) and NH3_cat.__init__ (OML/NH3_cat.py:20-83).            */
int so_lattice_create(int d, int kind, int device, so_lattice **out);
int so_lattice_destroy(so_lattice *lat);
/* n_sites = d*d variable sites (layer 3), n_nodes = 5*d*d graph nodes, nnz = directed neighbour entries */
int so_lattice_info(const so_lattice *lat, int *d, int *n_sites, int *n_nodes, int *nnz, int *words);
/* neighbour CSR of the 5N-node graph (what ASE NeighborList + nx.Graph hold upstream, LSC_cat.py:71-79) */
int so_lattice_csr(const so_lattice *lat, int32_t *rowptr, int32_t *col);
/* raw cartesian positions [n_nodes][3] and the 2x2 in-plane cell (atoms_template.get_positions/get_cell) */
int so_lattice_positions(const so_lattice *lat, double *pos, double *cell);

/* ---- chains: the occupancy state (dynamic_cat.variable_occs, OML/dynamic_cat.py:36-37,57-58) for
 *      n_chains independent structures, bit-packed (ceil(N/32) words per chain, site v = bit v&31 of word v>>5) */
int so_chains_create(so_lattice *lat, int64_t n_chains, so_chains **out);
int so_chains_destroy(so_chains *c);
/* assign_occs (OML/dynamic_cat.py:152-159): occ is [n][N] bytes in {0,1}; any other value -> SO_EINVAL   */
int so_set_occupancy(so_chains *c, int64_t chain0, int64_t n, const uint8_t *occ);
int so_get_occupancy(so_chains *c, int64_t chain0, int64_t n, uint8_t *occ);
int so_set_occupancy_bits(so_chains *c, int64_t chain0, int64_t n, const uint32_t *bits);
int so_get_occupancy_bits(so_chains *c, int64_t chain0, int64_t n, uint32_t *bits);
/* randomize (OML/dynamic_cat.py:62-84) on device: Bernoulli(p) bits per site, p = coverage, or p ~ U(0,1)
 * per chain when coverage < 0 (the reference's coverage=None); Philox stream `seed`.                      */
int so_randomize(so_chains *c, uint64_t seed, double coverage);
/* seeded island structures, the initial-population recipe of drivers/generate_init_structures.py:26-116, on the device:
 * chain chain0+k becomes structure index[k] of n_strucs (np.random.seed(index[k]); int(3 + 15*i/(n_strucs+1)) 19-site
 * hexagonal islands around sites drawn without replacement; then n_mutate (16 upstream) sites toggled) -- NumPy's
 * legacy MT19937 / choice(replace=False) stream restated bit for bit.                                        */
int so_island_structures(so_chains *c, int64_t chain0, int64_t n, const int32_t *index, int n_strucs, int n_mutate);
/* flip_atom (OML/dynamic_cat.py:172-185) for a list of (chain, variable-site index) pairs; keeps the
 * surrogate state (if attached) consistent.                                                             */
int so_flip(so_chains *c, int64_t n, const int64_t *chains, const int32_t *sites);

/* ---- descriptors / site types: replaces the Ni-neighbour count pass (OML/LSC_cat.py:186-196) and
 *      graph_to_KMClattice (OML/LSC_cat.py:164-245, OML/NH3_cat.py:158-428).  Any output may be NULL.
 *      c6        [n][N]   Ni count among the 6 in-plane neighbours
 *      lsc_type  [n][N]   0 (None) / 1 / 2
 *      nh3_type  [n][5N]  0 (None) / 1..19 for every graph node
 *      lsc_hist  [n][3]   counts of lsc_type 0,1,2
 *      nh3_hist  [n][20]  counts of the 19 types (column 0 = None)
 *      nh3_exp   [n][8]   counts of the exported KMC site-type index 1..7 (NH3_cat.py:421; column 0 = rest) */
int so_descriptors(so_chains *c, int64_t chain0, int64_t n, uint8_t *c6, uint8_t *lsc_type,
                   uint8_t *nh3_type, int32_t *lsc_hist, int32_t *nh3_hist, int32_t *nh3_exp);
/* per-flip descriptor deltas for one proposed site per chain (sites[n], chains chain0..chain0+n-1), WITHOUT
 * applying the flip: d_lsc[n][3], d_nh3_exp[n][8] = histogram(after) - histogram(before).  This is the
 * incremental typing NH3_cat.flip_atom (OML/NH3_cat.py:115-155) was meant to provide.                     */
int so_flip_delta(so_chains *c, int64_t chain0, int64_t n, const int32_t *sites, int32_t *d_lsc,
                  int32_t *d_nh3_exp);

/* ---- surrogate: the evaluate side of OML/train_surrogate.py:34-54.
 *      offsets [K][2]  lattice offset (o1,o2) of descriptor column k: row r, column k = x[r + offsets[k]]
 *                      (full translation row, OML/dynamic_cat.py:233-267: offsets[k] = (k % d, k / d);
 *                       7x7 local window, OML/dynamic_cat.py:325-341: d2 outer, d1 inner in [-3,3])
 *      W1 [K][H], b1 [H], W2 [H], b2: MLPRegressor coefs_/intercepts_ (relu hidden, identity out)
 *      y_mean, y_scale: the StandardScaler on Y (train_surrogate.py:141-144)
 *      tree: flattened DecisionTreeClassifier gate (train_surrogate.py:40,111-114) or n_tree_nodes = 0 for
 *            "all sites active": node i is a leaf iff feature[i] < 0; go left iff x[feature] <= thr[i];
 *            leaf_active[i] = 1 iff the leaf's majority class == 1.
 *      The device tables are exact fixed-point images of the weights (DESIGN.md "Arithmetic").           */
int so_surrogate_create(so_lattice *lat, int K, int H, const int32_t *offsets, const double *W1,
                        const double *b1, const double *W2, double b2, double y_mean, double y_scale,
                        int n_tree_nodes, const int32_t *feature, const double *thr, const int32_t *left,
                        const int32_t *right, const uint8_t *leaf_active, so_surrogate **out);
int so_surrogate_destroy(so_surrogate *s);
/* the same surrogate with a COMPACT first-layer state: the power-of-two grid of the first layer is 2^8 coarser
 * (|pre-activation| < 2^23 grid units), so that a site's 20 pre-activations pack into 64 bytes (five 24-bit values per
 * 16 bytes) instead of 80.  Records then coincide with the 64-byte fill granule of the L2: a flip moves 3,136 bytes of
 * DRAM traffic instead of ~4,480.  Scores stay within the 1e-5 bar (tests); decisions are those of the same exact
 * integer arithmetic on the coarser grid, re-taken from the fp64 weights near ties when exact_guard is set.  Needs 17..20
 * hidden units and a rectangular-window descriptor (e.g. get_local_inds); so_anneal then always runs the
 * pipelined window kernel.                                                                                       */
int so_surrogate_create_compact(so_lattice *lat, int K, int H, const int32_t *offsets, const double *W1,
                                const double *b1, const double *W2, double b2, double y_mean, double y_scale,
                                so_surrogate **out);
/* the compact state WITH the classifier gate of OML/train_surrogate.py:34-54 (same tree arrays as so_surrogate_create):
 * so_anneal then runs the gated window kernel, whose per-row gate records (16 bytes per row, allocated by
 * so_chains_attach) travel through the same TMA pipeline as the state.                                          */
int so_surrogate_create_compact_gated(so_lattice *lat, int K, int H, const int32_t *offsets, const double *W1,
                                      const double *b1, const double *W2, double b2, double y_mean, double y_scale,
                                      int n_tree_nodes, const int32_t *feature, const double *thr, const int32_t *left,
                                      const int32_t *right, const uint8_t *leaf_active, so_surrogate **out);
/* quantisation exponents chosen at load: h grid 2^e_h, w2 grid 2^e_w, padded hidden width Hp              */
int so_surrogate_info(const so_surrogate *s, int *K, int *H, int *Hp, int *e_h, int *e_w, int *gated);

/* attach: allocates the per-chain first-layer state hq[n_chains][N][Hp] (int32) and the objective sums      */
int so_chains_attach(so_chains *c, so_surrogate *s);
/* (re)compute state + objective of chains [chain0, chain0+n) from their bits; of_out[n] may be NULL.
 * surrogate.eval_rate(cat.generate_all_translations()) upstream (optimize_SA.py:41-42).                   */
int so_score(so_chains *c, int64_t chain0, int64_t n, double *of_out);
/* current tracked objective (no recompute); sums as exact integers: hi*2^20+lo, n_active; any may be NULL */
int so_objective(so_chains *c, int64_t chain0, int64_t n, double *of_out, int64_t *hi, int64_t *lo,
                 int32_t *n_active);
/* debug/parity: copy the first-layer state of one chain, hq[N][Hp] int32                                 */
int so_get_preact(so_chains *c, int64_t chain, int32_t *hq);
/* stateless batch scoring (eval_init_structures analogue, BASELINE config 3): bits[n][words] ->
 * rate[n] (+ optional histograms as in so_descriptors)                                                   */
int so_score_batch(so_lattice *lat, so_surrogate *s, int64_t n, const uint32_t *bits, double *rate,
                   int32_t *lsc_hist, int32_t *nh3_exp);
/* generic forward over arbitrary descriptor rows X[n_rows][K] (bytes 0/1): per-row active flag and rate,
 * i.e. classifier.predict + predictor.predict + inverse_transform (train_surrogate.py:40-50)             */
int so_eval_rows(so_surrogate *s, int64_t n_rows, const uint8_t *X, uint8_t *active, double *rate);

/* ---- annealing: the hot loop of optimize() (OML/optimize_SA.py:59-132).
 *      Runs steps [step0, step0+n_steps) of every chain:
 *        T = tscale[chain] * temps[s - step0]              (temps computed by the caller, optimize_SA.py:61-68)
 *        f = proposals ? proposals[chain][s-step0] : mulhi(philox(seed; chain_id, s)[0], N)
 *        u = uniforms  ? uniforms[chain][s-step0]  : (philox(...)[1] >> 8) * 2^-24
 *        delta = OF(x ^ e_f) - OF(x)                        (exact integer sums -> fp64)
 *        accept iff delta > 0, or T > 0 and exp(delta / T) > u      (optimize_SA.py:108-116)
 *      proposals/uniforms: [n_chains][n_steps] host arrays or NULL.  accept_out: [n_chains][n_steps] bytes
 *      or NULL.  chain_id for the Philox counter = chain_id0 + local chain index.                        */
typedef struct so_anneal_args {
  int64_t step0;
  int64_t n_steps;
  const double *temps;       /* [n_steps] host                                     */
  uint64_t seed;
  int64_t chain_id0;         /* global id of local chain 0 (multi-GPU sharding)    */
  const int32_t *proposals;  /* optional injected proposal sequence                */
  const float *uniforms;     /* optional injected uniform sequence                 */
  uint8_t *accept_out;       /* optional accept/reject log                         */
  int exact_guard;           /* 1: re-decide near-ties with fp64 weights           */
  double guard_tol;          /* |delta - T ln u| below which the guard fires       */
  int variant;               /* 0 = auto; tests: 1 = row-per-lane global-memory kernel, 2 = generic flat kernel,
                                3 = auto without the shared-memory-resident kernels of small lattices,
                                4 = resident warp-per-chain kernel, 5 = resident CTA-per-chain kernel,
                                6 = cached-gate kernel with global state (instead of the gated window kernel) */
} so_anneal_args;

typedef struct so_anneal_stats {
  int64_t flips;             /* chain-steps evaluated                              */
  int64_t accepted;
  int64_t guard_fired;
  float kernel_ms;           /* CUDA-event time of the annealing kernel(s) only     */
  int kernel_id;             /* which kernel the dispatch chose (so_kernel_name)    */
} so_anneal_stats;

int so_anneal(so_chains *c, const so_anneal_args *args, so_anneal_stats *stats);
/* name of the kernel behind so_anneal_stats.kernel_id ("sa_window_kernel", "sa_resident_kernel", ...)   */
const char *so_kernel_name(int kernel_id);
/* per-chain temperature scale (parallel-tempering rung), default 1.0                                      */
int so_set_tscale(so_chains *c, int64_t chain0, int64_t n, const double *tscale);
int so_get_tscale(so_chains *c, int64_t chain0, int64_t n, double *tscale);

/* ---- site-type histograms carried through the annealing.  The reference commits "x, syms, OF" per accepted step
 *      (OML/optimize_SA.py:119-123) and re-types the final structure (drivers/Online_learn_main.py:190); its NH3_cat.flip_atom
 *      (OML/NH3_cat.py:115-155) was meant to re-type incrementally.  so_track_types(c, 1) makes every SA kernel apply the
 *      per-flip histogram delta (the rule of so_flip_delta) inside the accept path, so that the LSC type counts {None, 1, 2}
 *      and the exported NH3 type counts {not exported, 1..7} of every chain are always those of its current structure;
 *      so_type_histograms reads them (lsc_hist[n][3], nh3_exp[n][8], either may be NULL) without re-typing.
 *      Any other change of the bits (set_occupancy, flip, randomize, load) recomputes them.  Off by default.        */
int so_track_types(so_chains *c, int enable);
int so_type_histograms(so_chains *c, int64_t chain0, int64_t n, int32_t *lsc_hist, int32_t *nh3_exp);

/* ---- checkpoint / resume.  The reference's SA is not resumable (its "checkpoint" is the database folder,
 *      OML/KMC_handler.py:37-42); here a run continues bit for bit from (bits, tscale, step, seed) because the draws are
 *      counter based and the first-layer state is an exact function of the bits.  so_chains_save writes bits + tscale +
 *      the exact objective sums of all chains with the caller's step counter and seed; so_chains_load restores them
 *      into a handle of the same shape, rebuilds the state and verifies the sums (SO_EINVAL: wrong shape / not a
 *      checkpoint / written with another surrogate).  step / seed may be NULL.                                */
int so_chains_save(so_chains *c, const char *path, int64_t step, uint64_t seed);
int so_chains_load(so_chains *c, const char *path, int64_t *step, uint64_t *seed);

/* ---- analytical objective of the LSC toy model.
 *      so_lsc_analytical: LSC_cat.compute_site_rates_lsc / eval_struc_rate_anal (OML/LSC_cat.py:248-316) of the current
 *      occupancies of chains [chain0, chain0+n): site_rates[n][N] and/or struct_rate[n] (either may be NULL).
 *      so_anneal_analytical: the loop of optimize(mode='analytical') (OML/optimize_SA.py:59-132, objective :104) with
 *      the same arguments as so_anneal (exact_guard / variant ignored); of_out[n_chains] = final objective or NULL.
 *      The dense steady-state system lives in shared memory: lattices up to 160 sites (d <= 12).              */
int so_lsc_analytical(so_chains *c, int64_t chain0, int64_t n, double *site_rates, double *struct_rate);
int so_anneal_analytical(so_chains *c, const so_anneal_args *args, so_anneal_stats *stats, double *of_out);

/* ---- best structure + replica exchange (new capability; not in the reference)               */
/* arg-max of the tracked objective over local chains; bits[words] may be NULL                            */
int so_best(so_chains *c, double *of_out, int64_t *chain_out, uint32_t *bits);
/* same, written as one record into DEVICE memory (e.g. a torch tensor used as NCCL send buffer) on `stream`:
 * record = { double of; int64 chain_id; uint32 bits[words] }  (16 + 4*words bytes)                       */
int so_best_record_dev(so_chains *c, int64_t chain_id0, void *record_dev, void *stream);
/* tracked objectives of all local chains into DEVICE memory of_dev[n_chains] (fp64) on `stream`          */
int so_objectives_dev(so_chains *c, double *of_dev, void *stream);
/* parallel-tempering swap round.  Global chain g = i*world + rank (local chain i of `rank`: ladders of
 * `ladder` consecutive global ids are strided across the ranks, so neighbouring rungs sit on different GPUs).
 * of_all_dev / tscale_all_dev are DEVICE arrays in all-gather (rank-major) layout [world][n_chains]; every rank
 * takes the same decisions (Philox counter = (ladder index, round, rung)) and applies them to its own chains:
 * rungs ranked by current temperature, adjacent rungs (r, r+1), r = parity (mod 2), swap TEMPERATURES with
 * probability min(1, exp[(OF_a - OF_b) * (1/T_b - 1/T_a)]), T = tscale * t_ref.  tscale_all_dev is updated in
 * place (identically on every rank); n_swapped counts local chains whose temperature changed.           */
int so_exchange(so_chains *c, int rank, int world, int ladder, int parity, uint64_t seed, int64_t round,
                double t_ref, const double *of_all_dev, double *tscale_all_dev, void *stream,
                int64_t *n_swapped);

/* ---- regressor training on the device (replaces MLPRegressor(...).fit of OML/train_surrogate.py:141-152) ----------
 * The K -> H -> 1 ReLU MLP fitted by mini-batch Adam exactly as scikit-learn runs it (squared loss / 2 +
 * alpha / 2 |coefs|^2 / batch; one Adam step per batch; stop when the epoch loss has not improved by tol for more than
 * n_iter_no_change epochs), fp64, optimiser state in the distributed shared memory of one persistent thread-block cluster of 8 CTAs.
 *   X [n_rows][K] 0/1 descriptor rows (anything else: SO_EINVAL), y [n_rows] scaled targets, W1 [K][H] / b1 [H] /
 *   w2 [H] / b2 initial parameters (e.g. sklearn's Glorot draws).  so_trainer_run() advances n_epochs epochs (or until
 *   the stop rule fires); orders [n_epochs][n_rows] is the sample order of each epoch (batches are consecutive slices
 *   of batch_size), loss_out [n_epochs] receives the epoch losses (0 for epochs not run).                          */
typedef struct so_trainer so_trainer;
typedef struct so_train_params {
  double lr, alpha, beta1, beta2, eps, tol; /* sklearn: learning_rate_init, alpha, beta_1, beta_2, epsilon, tol */
  int batch_size;                           /* sklearn 'auto' = min(200, n_rows); clipped to n_rows              */
  int n_iter_no_change;                     /* sklearn default 10                                                */
} so_train_params;
typedef struct so_train_status {
  int epochs_done;      /* epochs run by this call                     */
  int n_iter;           /* epochs run in total (sklearn n_iter_)        */
  int stopped;          /* the tol rule has fired                       */
  int64_t adam_steps;   /* optimiser steps so far (sklearn _optimizer.t) */
  double best_loss;
  float kernel_ms;
} so_train_status;
int so_trainer_create(int device, int64_t n_rows, int K, int H, const uint8_t *X, const double *y, const double *W1,
                      const double *b1, const double *w2, double b2, const so_train_params *hp, so_trainer **out);
int so_trainer_destroy(so_trainer *t);
int so_trainer_run(so_trainer *t, int n_epochs, const int32_t *orders, double *loss_out, so_train_status *status);
int so_trainer_get(so_trainer *t, double *W1, double *b1, double *w2, double *b2);
/* MLPRegressor.predict with the current parameters: X [n][K] 0/1 rows -> out [n] (scaled units, fp64) */
int so_trainer_predict(so_trainer *t, int64_t n, const uint8_t *X, double *out);

/* ---- classifier training on the device (replaces DecisionTreeClassifier(max_depth=30, random_state=0).fit of
 *      OML/train_surrogate.py:111-114) and the training-set construction before it (:72-98 with the x3 rotation
 *      augmentation of OML/dynamic_cat.py:270-322).
 * so_tree_fit: scikit-learn's CART for a 0/1 matrix X[n_rows][K] and class indices y[n_rows] in {0, 1} (two_classes = 0:
 *   all labels equal), Gini, best splitter, depth-first build, feature order from scikit-learn's xorshift stream seeded
 *   with splitter_seed (= RandomState(random_state).randint(0, 2^31 - 1)), so that the result is node for node the tree
 *   scikit-learn fits.  Outputs (capacity max_nodes, SO_EINVAL if exceeded) in scikit-learn's layout: feature (-2 = leaf),
 *   threshold (0.5 / -2), children_left / children_right (-1 = none), class counts per node; *n_nodes nodes.
 * so_build_training_set: for n_struct structures (bits[n_struct][words]) the 3N x N matrix of all translations and
 *   rotations, rows 3*(s*N + r) + k, as bytes into X[n_struct*3N][N] (may be NULL); with site_rates[n_struct][N]: the
 *   maximum rate (*y_max) and the activity label of every row, active[n_struct*3N] = rate > frac * max (frac = 0.06
 *   upstream); X_reg / Y_reg of the reference are the rows / tiled rates with active = 1.                      */
int so_tree_fit(int device, int64_t n_rows, int K, const uint8_t *X, const uint8_t *y, int two_classes, int max_depth,
                uint32_t splitter_seed, int max_nodes, int32_t *feature, double *threshold, int32_t *children_left,
                int32_t *children_right, double *count0, double *count1, int32_t *n_nodes, float *kernel_ms);
int so_build_training_set(so_lattice *lat, int64_t n_struct, const uint32_t *bits, const double *site_rates, double frac,
                          uint8_t *X, uint8_t *active, double *y_max);

#ifdef __cplusplus
}
#endif
#endif /* SO_B200_H */
