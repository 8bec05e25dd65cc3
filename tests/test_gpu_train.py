"""Device-side regressor training (so_train.cu) against the training oracle and scikit-learn itself
(OML/train_surrogate.py:141-152).  fp64 on both sides; only the summation order differs, tolerances are stated."""
import warnings
import numpy as np
import pytest
from oracle import train as T
from oracle import lattice as L, lsc_analytical as LA

pytestmark = pytest.mark.gpu


def _data(n, K, seed):
    rng = np.random.default_rng(seed)
    X = (rng.random((n, K)) < 0.5).astype(np.float64)
    w = rng.normal(size=K) / np.sqrt(K)
    y = np.maximum(X @ w, 0) + 0.05 * rng.normal(size=n)
    return X, (y - y.mean()) / y.std()


def _close(dev, ref_W1, ref_b1, ref_w2, ref_b2, rtol, atol):
    np.testing.assert_allclose(dev.coefs_[0], ref_W1, rtol=rtol, atol=atol)
    np.testing.assert_allclose(dev.intercepts_[0], ref_b1, rtol=rtol, atol=atol)
    np.testing.assert_allclose(dev.coefs_[1].ravel(), np.ravel(ref_w2), rtol=rtol, atol=atol)
    np.testing.assert_allclose(dev.intercepts_[1][0], ref_b2, rtol=rtol, atol=atol)


@pytest.mark.parametrize("n,K,max_iter", [(950, 49, 40), (333, 144, 25), (57, 16, 60), (1, 5, 3), (201, 256, 6)])
def test_device_fit_matches_oracle(n, K, max_iter):
    from so_b200.train_surrogate import DeviceMLPRegressor
    X, y = _data(n, K, seed=n) if n > 1 else (np.ones((1, K)), np.array([0.3]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dev = DeviceMLPRegressor(learning_rate_init=0.0001, alpha=0.1, max_iter=max_iter, tol=0.00001, random_state=7,
                                 epochs_per_launch=7).fit(X, y)
    o = T.fit_adam(X, y, random_state=7, max_iter=max_iter)
    assert dev.n_iter_ == o["n_iter"]
    np.testing.assert_allclose(dev.loss_curve_, o["loss_curve"], rtol=1e-10)          # fp64, different summation order
    _close(dev, o["W1"], o["b1"], o["w2"], o["b2"], rtol=1e-8, atol=1e-12)


def test_device_fit_matches_sklearn_and_stops_alike():
    """The reference's own call (MLPRegressor.fit) with a learning rate / tol at which the stop rule fires."""
    from sklearn.neural_network import MLPRegressor
    from so_b200.train_surrogate import DeviceMLPRegressor
    X, y = _data(400, 30, seed=5)
    kw = dict(activation='relu', learning_rate_init=0.01, alpha=0.1, hidden_layer_sizes=(20,), max_iter=2000, tol=1e-3)
    rs_ref, rs_dev = np.random.RandomState(3), np.random.RandomState(3)
    ref = MLPRegressor(random_state=rs_ref, **kw).fit(X, y)
    dev = DeviceMLPRegressor(random_state=rs_dev, **kw).fit(X, y)
    # the generator is left where scikit-learn leaves it (the stop fires inside a chunk of prepared epoch orders)
    assert ref.n_iter_ % 16 != 0 and rs_ref.randint(1 << 30) == rs_dev.randint(1 << 30)
    assert ref.n_iter_ < 2000 and dev.n_iter_ == ref.n_iter_
    np.testing.assert_allclose(dev.loss_curve_, ref.loss_curve_, rtol=1e-8)
    _close(dev, ref.coefs_[0], ref.intercepts_[0], ref.coefs_[1], ref.intercepts_[1][0], rtol=1e-6, atol=1e-10)
    np.testing.assert_allclose(dev.predict(X[:50]), ref.predict(X[:50]), rtol=1e-6, atol=1e-9)
    assert abs(dev.best_loss_ - ref.best_loss_) <= 1e-8 * ref.best_loss_


def test_surrogate_train_on_device_matches_host():
    """surrogate.train(): GPU-fitted regressor == scikit-learn-fitted regressor on the reference recipe (d = 12, three
    rotations per translation row, 0.06 * max activity threshold), and the structure rates they give agree."""
    from so_b200 import LSC_cat, surrogate
    d = 12
    occs = np.array([LA.make_random_structure(i, 48, d) for i in range(0, 48, 6)])
    rates = np.array([LA.compute_site_rates_lsc(x, d) for x in occs])
    all_syms = np.vstack([L.generate_all_translations_and_rotations(x, d) for x in occs])
    out = []
    for on_device in (True, False):
        sur = surrogate()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sur.train(all_syms, rates, max_iter=40, random_state=0, on_device=on_device)
        sur.cat = LSC_cat(d)
        out.append((sur, [sur.eval_rate(L.generate_all_translations(x, d)) for x in occs[:3]]))
    (a, ra), (b, rb) = out
    assert type(a.predictor).__name__ == "DeviceMLPRegressor" and type(b.predictor).__name__ == "MLPRegressor"
    np.testing.assert_allclose(a.predictor.loss_curve_, b.predictor.loss_curve_, rtol=1e-9)
    np.testing.assert_allclose(a.predictor.coefs_[0], b.predictor.coefs_[0], rtol=1e-7, atol=1e-11)
    np.testing.assert_allclose(ra, rb, rtol=1e-6)
    assert abs(a.mean_abs_error - b.mean_abs_error) <= 1e-6 * b.mean_abs_error


def test_trainer_rejects_bad_input():
    from so_b200 import engine
    X, y = _data(50, 12, seed=1)
    W1, b1, w2, b2 = T.init_params(np.random.RandomState(0), 12, 20)
    with pytest.raises(ValueError):
        engine.Trainer(X * 0.5, y, W1, b1, w2, b2)                      # not 0/1
    with pytest.raises(ValueError):
        engine.Trainer(X, y[:-1], W1, b1, w2, b2)
    Wbig = np.zeros((3000, 20))
    with pytest.raises(ValueError):
        engine.Trainer(np.zeros((10, 3000)), np.zeros(10), Wbig, b1, w2, b2)   # too wide for shared memory
    tr = engine.Trainer(X, y, W1, b1, w2, b2)
    with pytest.raises(ValueError):
        tr.run(np.full((1, 50), 50, dtype=np.int32))                    # order index out of range
    loss, st = tr.run(np.zeros((0, 50), dtype=np.int32))
    assert len(loss) == 0 and st["n_iter"] == 0
    tr.close()
