"""The training oracle (oracle/train.py) pinned against the installed scikit-learn: same random_state => the same
initial weights, epoch orders, loss curve and fitted weights (OML/train_surrogate.py:141-152 recipe)."""
import warnings
import numpy as np
import pytest
from oracle import train as T


def _data(n, K, seed):
    rng = np.random.default_rng(seed)
    X = (rng.random((n, K)) < 0.5).astype(np.float64)
    w = rng.normal(size=K) / np.sqrt(K)
    y = np.maximum(X @ w, 0) + 0.05 * rng.normal(size=n)
    return X, (y - y.mean()) / y.std()


@pytest.mark.parametrize("n,K,max_iter", [(950, 49, 40), (333, 144, 25), (57, 16, 60)])
def test_fit_adam_matches_sklearn(n, K, max_iter):
    from sklearn.neural_network import MLPRegressor
    X, y = _data(n, K, seed=n)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = MLPRegressor(activation='relu', learning_rate_init=0.0001, alpha=0.1, hidden_layer_sizes=(20,),
                           max_iter=max_iter, tol=0.00001, random_state=7).fit(X, y)
    o = T.fit_adam(X, y, random_state=7, max_iter=max_iter)
    assert o["n_iter"] == ref.n_iter_
    np.testing.assert_allclose(o["loss_curve"], ref.loss_curve_, rtol=1e-10)
    np.testing.assert_allclose(o["W1"], ref.coefs_[0], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(o["w2"], ref.coefs_[1].ravel(), rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(o["b1"], ref.intercepts_[0], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(o["b2"], ref.intercepts_[1][0], rtol=1e-8, atol=1e-12)


def test_fit_adam_stops_like_sklearn():
    """tol-based stop (n_iter_no_change = 10) at a learning rate that converges quickly."""
    from sklearn.neural_network import MLPRegressor
    X, y = _data(400, 30, seed=5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = MLPRegressor(activation='relu', learning_rate_init=0.01, alpha=0.1, hidden_layer_sizes=(20,),
                           max_iter=2000, tol=1e-3, random_state=3).fit(X, y)
    o = T.fit_adam(X, y, random_state=3, lr=0.01, tol=1e-3, max_iter=2000)
    assert ref.n_iter_ < 2000 and o["n_iter"] == ref.n_iter_
    np.testing.assert_allclose(o["loss_curve"], ref.loss_curve_, rtol=1e-8)


def test_orders_and_init_can_be_injected():
    X, y = _data(300, 20, seed=9)
    rng = np.random.RandomState(11)
    init = T.init_params(rng, 20, 20)
    orders, _ = T.epoch_orders(rng, 300, 12)
    a = T.fit_adam(X, y, random_state=11, max_iter=12)
    b = T.fit_adam(X, y, init=init, orders=orders, max_iter=12)
    assert (a["W1"] == b["W1"]).all() and (a["loss_curve"] == b["loss_curve"]).all()
