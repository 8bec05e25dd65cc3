"""Regressor training: device (so_train.cu) against scikit-learn on the host cores, reference recipe shapes
(d=12: rows = structures x 144 sites x 3 rotations, K = 144; and the 7x7 local descriptor K = 49)."""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "structure-optimization_b200"))
import numpy as np
from sklearn.neural_network import MLPRegressor
from so_b200.train_surrogate import DeviceMLPRegressor

warnings.simplefilter("ignore")
for n, K, epochs in ((20736, 144, 30), (20736, 49, 30), (2000, 144, 100)):
    rng = np.random.default_rng(0)
    X = (rng.random((n, K)) < 0.5).astype(np.float64)
    y = np.maximum(X @ (rng.normal(size=K) / np.sqrt(K)), 0) + 0.05 * rng.normal(size=n)
    y = (y - y.mean()) / y.std()
    kw = dict(activation='relu', learning_rate_init=0.0001, alpha=0.1, hidden_layer_sizes=(20,), max_iter=epochs, tol=1e-5,
              random_state=0)
    DeviceMLPRegressor(**kw).fit(X[:400], y[:400])          # warm-up (context, module load)
    t0 = time.perf_counter(); dev = DeviceMLPRegressor(**kw).fit(X, y); t_dev = time.perf_counter() - t0
    t0 = time.perf_counter(); ref = MLPRegressor(**kw).fit(X, y); t_ref = time.perf_counter() - t0
    nb = -(-n // 200) * dev.n_iter_
    print("n=%d K=%d: %d epochs; device %.3f s wall (kernel %.1f ms, %.2f us per Adam step), sklearn %.3f s (%.1f us per step); "
          "final loss %.6f vs %.6f" % (n, K, dev.n_iter_, t_dev, dev.kernel_ms_, 1e3 * dev.kernel_ms_ / nb, t_ref,
                                       1e6 * t_ref / nb, dev.loss_curve_[-1], ref.loss_curve_[-1]), flush=True)
