"""
surrogate -- drop-in for OML/train_surrogate.py:15-184.  Training runs on the GPU by default: the activity labels
(so_build_training_set), the tree classifier (DeviceTreeClassifier: scikit-learn's CART restated in so_tree.cu) and the
regressor (DeviceMLPRegressor: scikit-learn's Adam recipe restated in so_train.cu); `on_device=False` calls
scikit-learn itself for both fits.  The 75/25 split permutation and the Y scaler (a mean and a standard deviation) stay
on the host.  The evaluate side (eval_structure_rate, called `eval_rate` by optimize()) runs on the GPU.
Upstream defects resolved as in SURVEY 9.5: `Y_scaler` is stored, `train` takes self, both method names exist.
"""
import numpy as np
from . import engine


def descriptor_offsets(K, d, descriptor=None):
    """Site offsets of the K descriptor columns: full translation row (K == N, generate_all_translations order) or the 7x7
    local window (K == 49, get_local_inds order).  On a 7x7 lattice both have 49 columns in DIFFERENT orders: the caller
    must then say which one the surrogate was trained on (descriptor='full' | 'local')."""
    if descriptor not in (None, 'full', 'local'):
        raise ValueError("descriptor must be 'full' or 'local'")
    if K == d * d == 49 and descriptor is None:
        raise ValueError("49 columns on a 7x7 lattice: state descriptor='full' (translation row) or 'local' (7x7 window)")
    if descriptor == 'local' or (descriptor is None and K == 49 and K != d * d):
        if K != 49:
            raise ValueError('the local descriptor has 49 columns, the surrogate has %d' % K)
        return engine.local_offsets(3)
    if K == d * d:
        return engine.full_offsets(d)
    if K == 49:
        return engine.local_offsets(3)
    raise ValueError('descriptor width %d is neither the full translation row (%d) nor the 7x7 local window' % (K, d * d))


def flatten_tree(classifier):
    """sklearn DecisionTreeClassifier -> arrays for so_surrogate_create (leaf iff feature < 0; left iff x <= thr)."""
    t = classifier.tree_
    cls = np.asarray(classifier.classes_)
    leaf_cls = cls[np.argmax(t.value[:, 0, :], axis=1)]
    return dict(feature=t.feature.astype(np.int32), thr=t.threshold.astype(np.float64),
                left=t.children_left.astype(np.int32), right=t.children_right.astype(np.int32),
                active=(leaf_cls == 1.0).astype(np.uint8))


def tables_for(surr, lattice, gate=True, descriptor=None):
    """Device tables of any reference-style fitted surrogate (duck type: predictor, Y_scaler, classifier)."""
    mlp = surr.predictor
    if getattr(mlp, 'out_activation_', 'identity') != 'identity' or len(mlp.coefs_) != 2:
        raise ValueError('the surrogate must be a single-hidden-layer MLPRegressor (train_surrogate.py:150-151)')
    W1 = np.asarray(mlp.coefs_[0], dtype=np.float64)
    clf = getattr(surr, 'classifier', None) if gate else None
    scaler = getattr(surr, 'Y_scaler', None)
    y_mean = float(scaler.mean_[0]) if scaler is not None else 0.0
    y_scale = float(scaler.scale_[0]) if scaler is not None else 1.0
    return engine.SurrogateTables(lattice, descriptor_offsets(W1.shape[0], lattice.d, descriptor), W1, mlp.intercepts_[0],
                                  np.asarray(mlp.coefs_[1]).reshape(-1), float(np.asarray(mlp.intercepts_[1]).reshape(-1)[0]),
                                  y_mean, y_scale, flatten_tree(clf) if clf is not None else None)


class _TreeArrays(object):
    """The fitted-tree attributes of sklearn.tree._tree.Tree that the reference path and flatten_tree() read."""

    def __init__(self, feature, threshold, left, right, count0, count1):
        self.feature, self.threshold = feature, threshold
        self.children_left, self.children_right = left, right
        self.node_count = len(feature)
        tot = np.maximum(count0 + count1, 1.0)
        self.n_node_samples = (count0 + count1).astype(np.int64)
        self.weighted_n_node_samples = count0 + count1
        self.value = np.stack([count0 / tot, count1 / tot], axis=1)[:, None, :]      # class fractions, as sklearn >= 1.3


class DeviceTreeClassifier(object):
    """The DecisionTreeClassifier(max_depth=30, random_state=0) of OML/train_surrogate.py:111 fitted on the GPU
    (so_tree_fit): scikit-learn's CART on 0/1 descriptor rows -- Gini, best splitter, depth-first build, the feature
    order of scikit-learn's own xorshift stream seeded from `random_state` -- so the fitted tree is node for node the one
    `sklearn.tree.DecisionTreeClassifier(...).fit` returns.  Exposes classes_, tree_ (feature, threshold, children_left,
    children_right, value, node_count) and predict()."""

    def __init__(self, max_depth=30, random_state=None, device=0, max_nodes=None):
        self.max_depth, self.random_state, self._device, self._max_nodes = max_depth, random_state, device, max_nodes

    def fit(self, X, y):
        import ctypes as C
        from . import _lib as B
        X = np.asarray(X)
        Xu = np.ascontiguousarray(X != 0, dtype=np.uint8)
        if not np.array_equal(Xu, X):
            raise ValueError('training rows must be 0/1 occupancies')
        y = np.asarray(y)
        self.classes_ = np.unique(y)
        if len(self.classes_) > 2:
            raise ValueError('the device classifier fits the two-class activity gate of train_surrogate.py')
        yi = np.ascontiguousarray(np.searchsorted(self.classes_, y), dtype=np.uint8)
        n, K = Xu.shape
        rs = self.random_state
        rng = np.random.mtrand._rand if rs is None else (rs if isinstance(rs, np.random.RandomState) else np.random.RandomState(rs))
        seed = int(rng.randint(0, 2147483647))                  # Splitter.__cinit__: rand_r_state
        cap = int(self._max_nodes or min(2 * n + 1, 1 << 20))
        feature = np.empty(cap, dtype=np.int32)
        thr = np.empty(cap, dtype=np.float64)
        left = np.empty(cap, dtype=np.int32)
        right = np.empty(cap, dtype=np.int32)
        c0 = np.empty(cap, dtype=np.float64)
        c1 = np.empty(cap, dtype=np.float64)
        nn = C.c_int32()
        ms = C.c_float()
        depth = 62 if self.max_depth is None else int(self.max_depth)
        p = lambda a, t: a.ctypes.data_as(t)
        B.check(B.load().so_tree_fit(int(self._device), n, K, p(Xu, B.c_u8p), p(yi, B.c_u8p), int(len(self.classes_) == 2), depth,
                                     seed, cap, p(feature, B.c_i32p), p(thr, B.c_f64p), p(left, B.c_i32p), p(right, B.c_i32p),
                                     p(c0, B.c_f64p), p(c1, B.c_f64p), C.byref(nn), C.byref(ms)))
        k = nn.value
        self.kernel_ms_ = ms.value
        self.tree_ = _TreeArrays(feature[:k].astype(np.int64), thr[:k].copy(), left[:k].astype(np.int64),
                                 right[:k].astype(np.int64), c0[:k].copy(), c1[:k].copy())
        self.n_features_in_ = K
        return self

    def predict(self, X):
        t = self.tree_
        X = np.asarray(X)
        node = np.zeros(len(X), dtype=np.int64)
        live = t.feature[node] >= 0
        while live.any():
            idx = np.nonzero(live)[0]
            nd = node[idx]
            go_left = X[idx, t.feature[nd]] <= t.threshold[nd]
            node[idx] = np.where(go_left, t.children_left[nd], t.children_right[nd])
            live = t.feature[node] >= 0
        return self.classes_[np.argmax(t.value[node, 0, :len(self.classes_)], axis=1)]


def build_training_set(lattice, occupancies, site_rates=None, frac=0.06, want_rows=True):
    """all_syms (every translation x 3 rotations of every structure, OML/dynamic_cat.py:270-322), and -- with site rates --
    the activity label of every row and the maximum rate (OML/train_surrogate.py:72-98), computed on the device
    (so_build_training_set).  Returns (all_syms float64 [n*3N, N] or None, site_is_active float64 [n*3N] or None, y_max)."""
    import ctypes as C
    from . import _lib as B
    occ = np.asarray(occupancies)
    if occ.ndim == 1:
        occ = occ[None, :]
    n, N = occ.shape
    bits = engine.pack_bits(occ, lattice.W)
    X = np.empty((n * 3 * N, N), dtype=np.uint8) if want_rows else None
    rates = active = None
    ymax = C.c_double(0.0)
    if site_rates is not None:
        rates = np.ascontiguousarray(np.asarray(site_rates, dtype=np.float64).reshape(n, N))
        active = np.empty(n * 3 * N, dtype=np.uint8)
    p = lambda a, t: a.ctypes.data_as(t) if a is not None else None
    B.check(B.load().so_build_training_set(lattice.h, n, p(bits, B.c_u32p), p(rates, B.c_f64p), float(frac), p(X, B.c_u8p),
                                           p(active, B.c_u8p), C.byref(ymax)))
    return (X.astype(np.float64) if X is not None else None, active.astype(np.float64) if active is not None else None,
            ymax.value if rates is not None else None)


class DeviceMLPRegressor(object):
    """The MLPRegressor of OML/train_surrogate.py:150-151 fitted on the GPU (so_trainer_*): scikit-learn's mini-batch
    Adam with L2 penalty, its Glorot initialisation and its epoch shuffles drawn from the same `random_state`, so a
    fit reproduces `sklearn.neural_network.MLPRegressor(...).fit` to rounding.  Exposes the fitted attributes the
    reference and tables_for() read (coefs_, intercepts_, out_activation_, loss_curve_, n_iter_, best_loss_)."""

    out_activation_ = 'identity'

    def __init__(self, hidden_layer_sizes=(20,), activation='relu', learning_rate_init=0.001, alpha=0.0001,
                 max_iter=200, tol=1e-4, random_state=None, verbose=False, batch_size='auto', beta_1=0.9, beta_2=0.999,
                 epsilon=1e-8, n_iter_no_change=10, device=0, epochs_per_launch=16):
        if activation != 'relu' or len(tuple(hidden_layer_sizes)) != 1:
            raise ValueError('the device trainer fits the single-hidden-layer ReLU regressor of train_surrogate.py')
        self.hidden_layer_sizes = tuple(hidden_layer_sizes)
        self.learning_rate_init, self.alpha, self.max_iter, self.tol = learning_rate_init, alpha, max_iter, tol
        self.random_state, self.verbose, self.batch_size = random_state, verbose, batch_size
        self.beta_1, self.beta_2, self.epsilon, self.n_iter_no_change = beta_1, beta_2, epsilon, n_iter_no_change
        self._device, self._chunk = device, int(epochs_per_launch)
        self._trainer = None

    @staticmethod
    def _rng(random_state):
        if random_state is None:
            return np.random.mtrand._rand                  # sklearn.utils.check_random_state(None)
        if isinstance(random_state, np.random.RandomState):
            return random_state
        return np.random.RandomState(random_state)

    def fit(self, X, y):
        X = np.asarray(X)
        y = np.asarray(y, dtype=np.float64).reshape(-1)
        n, K = X.shape
        H = int(self.hidden_layer_sizes[0])
        rng = self._rng(self.random_state)
        # sklearn _init_coef: Glorot uniform, draws in the order coef 0, intercept 0, coef 1, intercept 1
        b = np.sqrt(6.0 / (K + H))
        W1 = rng.uniform(-b, b, (K, H))
        b1 = rng.uniform(-b, b, H)
        b = np.sqrt(6.0 / (H + 1))
        w2 = rng.uniform(-b, b, (H, 1)).reshape(-1)
        b2 = float(rng.uniform(-b, b, 1)[0])
        if self._trainer is not None:
            self._trainer.close()
        self._trainer = engine.Trainer(X, y, W1, b1, w2, b2, lr=self.learning_rate_init, alpha=self.alpha,
                                       beta1=self.beta_1, beta2=self.beta_2, eps=self.epsilon, tol=self.tol,
                                       batch_size=None if self.batch_size == 'auto' else self.batch_size,
                                       n_iter_no_change=self.n_iter_no_change, device=self._device)
        # Epoch orders are drawn on the host exactly as sklearn.utils.shuffle draws them (cumulative permutations from the
        # same RandomState).  The next chunk is drawn by a helper thread while the GPU runs the current one (the C call
        # releases the GIL); when the stop rule fires inside a chunk, the generator is rewound to the state it has after
        # the last epoch actually run, so it is left exactly where MLPRegressor.fit would leave it.
        from concurrent.futures import ThreadPoolExecutor

        def draw_chunk(sample_idx, m):
            state = rng.get_state()
            orders = np.empty((m, n), dtype=np.int32)
            for e in range(m):
                perm = np.arange(n)
                rng.shuffle(perm)
                sample_idx = sample_idx[perm]
                orders[e] = sample_idx
            return state, orders, sample_idx

        curve, status, self.kernel_ms_ = [], dict(stopped=False, n_iter=0, best_loss=np.inf), 0.0
        planned = 0                                         # epochs whose orders have been drawn
        with ThreadPoolExecutor(max_workers=1) as pool:
            m = min(self._chunk, self.max_iter)
            nxt = pool.submit(draw_chunk, np.arange(n), m) if m > 0 else None
            while nxt is not None:
                state, orders, sample_idx = nxt.result()
                planned += len(orders)
                m = min(self._chunk, self.max_iter - planned)
                nxt = pool.submit(draw_chunk, sample_idx, m) if m > 0 else None
                loss, status = self._trainer.run(orders)
                self.kernel_ms_ += status['kernel_ms']
                curve.extend(loss.tolist())
                if self.verbose:
                    for i, v in enumerate(loss):
                        print('Iteration %d, loss = %.8f' % (len(curve) - len(loss) + i + 1, v))
                if status['stopped']:
                    if nxt is not None:
                        nxt.result()                        # let the helper finish before the generator is rewound
                        nxt = None
                    rng.set_state(state)
                    for e in range(status['epochs_done']):
                        rng.shuffle(np.arange(n))
        self.loss_curve_, self.n_iter_ = curve, status['n_iter']
        self.loss_, self.best_loss_ = (curve[-1] if curve else np.inf), status['best_loss']
        self.t_ = self.n_iter_ * n
        W1, b1, w2, b2 = self._trainer.get()
        self.coefs_ = [W1, w2.reshape(-1, 1)]
        self.intercepts_ = [b1, np.array([b2])]
        self.n_layers_, self.n_outputs_ = 3, 1
        if not status['stopped']:
            import warnings
            try:
                from sklearn.exceptions import ConvergenceWarning as _Warn
            except Exception:
                _Warn = UserWarning
            warnings.warn("Stochastic Optimizer: Maximum iterations (%d) reached and the optimization hasn't "
                          "converged yet." % self.max_iter, _Warn)
        return self

    def predict(self, X):
        if self._trainer is None:
            raise RuntimeError('this DeviceMLPRegressor is not fitted yet')
        return self._trainer.predict(np.asarray(X))


class surrogate(object):
    """Surrogate model replacing KMC evaluation (OML/train_surrogate.py:15-31)."""

    def __init__(self, device=0):
        self.cat = None
        self.predictor = None
        self.classifier = None
        self.Y_scaler = None
        self.all_syms = None
        self.X_reg = None
        self.Y_reg = None
        self.site_is_active = None
        self._device = device
        self._tables = None
        self._tables_key = None
        self.descriptor = None              # 'full' | 'local': only needed where K alone is ambiguous (7x7 lattice)

    # ---- evaluate (GPU) -----------------------------------------------------------------------------------
    def _get_tables(self, K):
        """Device tables of the CURRENT predictor / classifier / scaler / catalyst (rebuilt when any of them is replaced)."""
        key = (id(self.predictor), id(self.classifier), id(self.Y_scaler), id(self.cat), K)
        if self._tables is None or self._tables_key != key:
            if self.cat is not None:
                lattice = self.cat._lattice
            else:
                d = int(round(np.sqrt(K)))
                lattice = engine.Lattice(d if d * d == K and d >= 7 else 12, "LSC", device=self._device)
            self._tables = tables_for(self, lattice, descriptor=self.descriptor)
            self._tables_key = key
        return self._tables

    def eval_structure_rate(self, all_trans, normalize=False):
        """OML/train_surrogate.py:34-54: gate, MLP, inverse Y scaling, sum over active rows."""
        all_trans = np.asarray(all_trans)
        active, rate = self._get_tables(all_trans.shape[1]).eval_rows(all_trans)
        if not active.any():
            return 0
        total = np.sum(rate[active])
        return total / len(all_trans) if normalize else total

    eval_rate = eval_structure_rate          # the name optimize() calls (OML/optimize_SA.py:42,102)

    def eval_site_rate(self, all_trans, normalize=False):
        active, rate = self._get_tables(np.asarray(all_trans).shape[1]).eval_rows(all_trans)
        return np.where(active, rate, 0.0)

    # ---- train (host, scikit-learn; OML/train_surrogate.py:61-152) --------------------------------------------
    def train(self, structure_occs, site_rates, curr_fldr=None, max_iter=10000, random_state=None, verbose=False,
              on_device=True):
        self.all_syms = np.asarray(structure_occs)
        self.partition_data_set(np.asarray(site_rates), on_device=on_device)
        self.train_classifier(on_device=on_device)
        self.train_regressor(max_iter=max_iter, random_state=random_state, verbose=verbose, on_device=on_device)
        self._tables = None

    def partition_data_set(self, site_rates, on_device=True):
        """OML/train_surrogate.py:72-98: rates tiled x3 (one copy per rotation), active iff rate > 0.06 * max.  on_device:
        maximum and labels from so_build_training_set (the same float64 comparison), else NumPy as upstream."""
        site_rates = np.asarray(site_rates, dtype=np.float64)
        flat = np.transpose(np.tile(site_rates.flatten(), [3, 1])).flatten()      # one copy per rotation
        if on_device and self.cat is not None and site_rates.ndim == 2 and site_rates.shape[1] == self.cat.atoms_per_layer:
            dummy = np.zeros((site_rates.shape[0], site_rates.shape[1]), dtype=np.uint8)
            _, active, y_max = build_training_set(self.cat._lattice, dummy, site_rates, want_rows=False)
        else:
            y_max = np.max(flat)
            active = (flat > 0.06 * y_max).astype(float)
        if y_max == 0.:
            raise NameError('No active sites in database.')
        self.site_is_active = active
        self.X_reg = self.all_syms[self.site_is_active == 1]
        self.Y_reg = flat[self.site_is_active == 1]

    def train_classifier(self, on_device=True):
        """OML/train_surrogate.py:101-121.  The 75/25 split is scikit-learn's (one RandomState(0) permutation); the fit runs
        on the GPU (DeviceTreeClassifier, node for node scikit-learn's tree) or, on_device=False, in scikit-learn itself."""
        from sklearn.model_selection import train_test_split
        if on_device:
            self.classifier = DeviceTreeClassifier(max_depth=30, random_state=0, device=self._device)
        else:
            from sklearn import tree
            self.classifier = tree.DecisionTreeClassifier(max_depth=30, random_state=0)
        X_train, X_test, y_train, y_test = train_test_split(self.all_syms, self.site_is_active, random_state=0)
        self.classifier.fit(X_train, y_train)
        self.frac_wrong_train = float(np.sum(np.abs(self.classifier.predict(X_train) - y_train))) / len(y_train)
        self.frac_wrong_test = float(np.sum(np.abs(self.classifier.predict(X_test) - y_test))) / len(y_test)

    def train_regressor(self, reg_parity_fname=None, max_iter=10000, random_state=None, verbose=False, on_device=True):
        """OML/train_surrogate.py:141-152.  on_device=True fits the same model with the same optimiser on the GPU
        (DeviceMLPRegressor); on_device=False calls scikit-learn itself, as the reference does."""
        from sklearn.preprocessing import StandardScaler
        self.Y_scaler = StandardScaler(with_mean=True, with_std=True)
        self.Y_scaler.fit(self.Y_reg.reshape(-1, 1))
        Y = self.Y_scaler.transform(self.Y_reg.reshape(-1, 1)).flatten()
        if on_device:
            Regressor, extra = DeviceMLPRegressor, dict(device=self._device)
        else:
            from sklearn.neural_network import MLPRegressor as Regressor
            extra = {}
        self.predictor = Regressor(activation='relu', verbose=verbose, learning_rate_init=0.0001, alpha=0.1,
                                   hidden_layer_sizes=(20,), max_iter=max_iter, tol=0.00001,
                                   random_state=random_state, **extra)
        self.predictor.fit(self.X_reg, Y)
        pred = self.Y_scaler.inverse_transform(self.predictor.predict(self.X_reg).reshape(-1, 1)).flatten()
        self.mean_abs_error = float(np.mean(np.abs(pred - self.Y_reg)))
