"""
oracle/tree.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement of the classifier fit the reference runs (OML/train_surrogate.py:101-121):
    X_train, X_test, y_train, y_test = train_test_split(all_syms, site_is_active, random_state=0)
    DecisionTreeClassifier(max_depth=30, random_state=0).fit(X_train, y_train)
The arithmetic lives in scikit-learn (not vendored in the reference, un-pinned: README.rst:9); restated here from its
published algorithm for 0/1 feature matrices -- `sklearn/tree/_tree.pyx` DepthFirstTreeBuilder (LIFO stack: right child
pushed first, left child built first; node ids in build order), `_splitter.pyx` node_split_best (Fisher-Yates feature
draw with the known / found constant-feature bookkeeping; a candidate replaces the best only when its proxy
improvement is strictly larger), `_criterion.pyx` Gini (proxy = -n_R*gini_R - n_L*gini_L, evaluated in this order in
float64), `utils/_random.pxd` our_rand_r (32-bit xorshift 13/17/5, modulo 2^31) seeded with
RandomState(random_state).randint(0, 2^31 - 1) -- and PINNED against the installed scikit-learn itself
(tests/test_oracle_tree.py: identical node arrays on random and on reference-recipe data).

On a 0/1 matrix every non-constant feature has the single threshold 0.5; "left" = samples with the feature 0.
"""
import numpy as np

RAND_R_MAX = 2147483647
EPSILON = np.finfo("double").eps
DEFAULT_SEED = 1


class XorShift(object):
    def __init__(self, seed):
        self.s = int(seed) & 0xffffffff

    def rand_r(self):
        if self.s == 0:
            self.s = DEFAULT_SEED
        s = self.s
        s ^= (s << 13) & 0xffffffff
        s ^= s >> 17
        s ^= (s << 5) & 0xffffffff
        self.s = s
        return s % (RAND_R_MAX + 1)

    def rand_int(self, low, high):
        return low + self.rand_r() % (high - low)


def train_rows(n_samples, random_state=0, test_size=0.25):
    """Row indices of the training part of train_test_split(..., random_state=0): ShuffleSplit draws one permutation,
    the first ceil(0.25 n) entries are the test rows, the next floor(0.75 n) the training rows (in that order)."""
    n_test = int(np.ceil(test_size * n_samples))
    n_train = int(np.floor((1.0 - test_size) * n_samples))
    perm = np.random.RandomState(random_state).permutation(n_samples)
    return perm[n_test:n_test + n_train], perm[:n_test]


def _gini_children(l0, l1, r0, r1):
    wl, wr = l0 + l1, r0 + r1
    sq_l = 0.0 + l0 * l0
    sq_l += l1 * l1
    sq_r = 0.0 + r0 * r0
    sq_r += r1 * r1
    return 1.0 - sq_l / (wl * wl), 1.0 - sq_r / (wr * wr)


def fit_tree(X, y, max_depth=30, random_state=0):
    """DecisionTreeClassifier(max_depth, random_state).fit(X, y) for a 0/1 matrix X [n, K] and labels y in {0, 1}
    (floats as upstream).  Returns dict(feature, threshold, children_left, children_right, count0, count1, classes)."""
    X = np.asarray(X)
    Xb = X != 0
    y = np.asarray(y)
    classes = np.unique(y)
    yi = np.searchsorted(classes, y)                       # class index per sample
    n, K = Xb.shape
    rng = XorShift(np.random.RandomState(random_state).randint(0, RAND_R_MAX))
    features = list(range(K))
    constant_features = [0] * K
    feat, thr, left, right, c0, c1 = [], [], [], [], [], []
    two = len(classes) == 2
    # stack records: (sample index array, depth, parent, is_left, impurity, n_constant_features)
    stack = [(np.arange(n), 0, -1, 0, np.inf, 0)]
    first = True
    while stack:
        idx, depth, parent, is_left, impurity, n_known = stack.pop()
        n_node = len(idx)
        n1 = float((yi[idx] == 1).sum()) if two else 0.0
        n0 = float(n_node) - n1
        is_leaf = depth >= max_depth or n_node < 2
        if first:
            sq = 0.0 + n0 * n0
            sq += n1 * n1
            impurity = 1.0 - sq / (float(n_node) * float(n_node))
            first = False
        is_leaf = is_leaf or impurity <= EPSILON
        best = None
        n_total = n_known
        if not is_leaf:
            # ---- node_split_best ------------------------------------------------------------------------
            ones = Xb[idx].sum(0)                             # per feature: samples with the feature set
            ones1 = Xb[idx][yi[idx] == 1].sum(0) if two else np.zeros(K, dtype=np.int64)
            f_i, n_found, n_drawn, n_visited = K, 0, 0, 0
            best_proxy = -np.inf
            while f_i > n_total and (n_visited < K or n_visited <= n_found + n_drawn):
                n_visited += 1
                f_j = rng.rand_int(n_drawn, f_i - n_found)
                if f_j < n_known:
                    features[n_drawn], features[f_j] = features[f_j], features[n_drawn]
                    n_drawn += 1
                    continue
                f_j += n_found
                f = features[f_j]
                if ones[f] == 0 or ones[f] == n_node:        # constant in this node
                    features[f_j], features[n_total] = features[n_total], features[f_j]
                    n_found += 1
                    n_total += 1
                    continue
                f_i -= 1
                features[f_i], features[f_j] = features[f_j], features[f_i]
                r1 = float(ones1[f])
                r0 = float(ones[f]) - r1
                l1, l0 = n1 - r1, n0 - r0
                gl, gr = _gini_children(l0, l1, r0, r1)
                proxy = (-(r0 + r1) * gr - (l0 + l1) * gl)
                if proxy > best_proxy:
                    best_proxy = proxy
                    best = (f, gl, gr)
            features[:n_known] = constant_features[:n_known]
            constant_features[n_known:n_known + n_found] = features[n_known:n_known + n_found]
            is_leaf = best is None
        node_id = len(feat)
        if parent >= 0:
            (left if is_left else right)[parent] = node_id
        feat.append(-2 if is_leaf else best[0])
        thr.append(-2.0 if is_leaf else 0.5)
        left.append(-1)
        right.append(-1)
        c0.append(n0)
        c1.append(n1)
        if not is_leaf:
            f, gl, gr = best
            go_right = Xb[idx, f]
            stack.append((idx[go_right], depth + 1, node_id, 0, gr, n_total))
            stack.append((idx[~go_right], depth + 1, node_id, 1, gl, n_total))
    return dict(feature=np.array(feat, dtype=np.int64), threshold=np.array(thr), children_left=np.array(left, dtype=np.int64),
                children_right=np.array(right, dtype=np.int64), count0=np.array(c0), count1=np.array(c1), classes=classes)


def flat_arrays(t):
    """The arrays so_surrogate_create takes (leaf iff feature < 0; active = majority class is 1.0, ties -> class 0)."""
    if len(t["classes"]) == 2:
        leaf_cls = np.where(t["count1"] > t["count0"], t["classes"][1], t["classes"][0])
    else:
        leaf_cls = np.full(len(t["feature"]), t["classes"][0])
    return dict(feature=t["feature"].astype(np.int32), thr=t["threshold"].astype(np.float64),
                left=t["children_left"].astype(np.int32), right=t["children_right"].astype(np.int32),
                active=(leaf_cls == 1.0).astype(np.uint8))
