"""CPU: the CART restatement (oracle/tree.py) against the installed scikit-learn -- identical node arrays (feature,
threshold, children, class counts) on random 0/1 matrices and on reference-recipe data, and the train/test split."""
import os
import numpy as np
import pytest
from oracle import tree as T, lattice as L, surrogate as S
import helpers as Hp


def _same_tree(X, y, max_depth=30, random_state=0):
    from sklearn.tree import DecisionTreeClassifier
    clf = DecisionTreeClassifier(max_depth=max_depth, random_state=random_state).fit(X, y)
    t = T.fit_tree(X, y, max_depth=max_depth, random_state=random_state)
    sk = clf.tree_
    assert len(t["feature"]) == sk.node_count
    assert (t["feature"] == sk.feature).all() and (t["children_left"] == sk.children_left).all()
    assert (t["children_right"] == sk.children_right).all() and (t["threshold"] == sk.threshold).all()
    counts = sk.value[:, 0, :] * sk.weighted_n_node_samples[:, None]
    if len(clf.classes_) == 2:
        np.testing.assert_allclose(counts[:, 0], t["count0"], atol=1e-9)
        np.testing.assert_allclose(counts[:, 1], t["count1"], atol=1e-9)
    flat = T.flat_arrays(t)
    want = S.flatten_tree(clf)
    for k in want:
        assert (flat[k] == want[k]).all(), k
    return clf, t


@pytest.mark.parametrize("seed", range(12))
def test_random_binary_data(seed):
    rng = np.random.default_rng(seed)
    n, K = int(rng.integers(20, 1500)), int(rng.integers(3, 60))
    X = (rng.random((n, K)) < rng.random()).astype(np.float64)
    if seed % 3 == 0:
        X[:, rng.integers(0, K)] = 1.0                               # a constant column
        X[:, rng.integers(0, K)] = X[:, 0]                           # a duplicated column (exact ties)
    w = rng.normal(size=K)
    y = ((X @ w + 0.5 * rng.normal(size=n)) > np.median(X @ w)).astype(np.float64)
    _same_tree(X, y, max_depth=[30, 30, 4, 2][seed % 4], random_state=[0, 0, 7][seed % 3])


def test_degenerate_labels_and_tiny_sets():
    X = (np.random.default_rng(0).random((50, 8)) < 0.5).astype(np.float64)
    _same_tree(X, np.ones(50))                                       # one class: a single leaf, classes_ = [1.]
    _same_tree(X, np.zeros(50))
    _same_tree(X[:2], np.array([0.0, 1.0]))
    _same_tree(np.zeros((10, 4)), np.array([0.0, 1.0] * 5))          # no feature can split


def test_reference_recipe_tree_and_split():
    """The classifier of the reference-executed training run (tests/golden/ref_surrogate_sa.npz): the restated split +
    fit reproduce its flattened tree exactly."""
    from sklearn.model_selection import train_test_split
    z = np.load(os.path.join(Hp.GOLDEN, "ref_surrogate_sa.npz"))
    all_syms = np.vstack([L.generate_all_translations_and_rotations(x, 12) for x in z["train_occs"]])
    active = z["site_is_active"].astype(np.float64)
    tr, te = T.train_rows(len(active), random_state=0)
    X_train, X_test, y_train, y_test = train_test_split(all_syms, active, random_state=0)
    assert (all_syms[tr] == X_train).all() and (active[tr] == y_train).all() and (all_syms[te] == X_test).all()
    flat = T.flat_arrays(T.fit_tree(all_syms[tr], active[tr]))
    assert (flat["feature"] == z["t_feature"]).all() and (flat["left"] == z["t_left"]).all()
    assert (flat["right"] == z["t_right"]).all() and (flat["active"] == z["t_active"]).all()
    assert (flat["thr"] == z["t_thr"]).all()
