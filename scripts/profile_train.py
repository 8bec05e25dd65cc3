"""One chunk (8 epochs x 104 batches) of the device trainer at the reference recipe shape, for ncu."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "structure-optimization_b200"))
import numpy as np
from so_b200.train_surrogate import DeviceMLPRegressor
warnings.simplefilter("ignore")
n, K = 20736, 144
rng = np.random.default_rng(0)
X = (rng.random((n, K)) < 0.5).astype(np.float64)
y = np.maximum(X @ (rng.normal(size=K) / np.sqrt(K)), 0) + 0.05 * rng.normal(size=n)
y = (y - y.mean()) / y.std()
dev = DeviceMLPRegressor(learning_rate_init=0.0001, alpha=0.1, max_iter=8, tol=1e-5, random_state=0, epochs_per_launch=8).fit(X, y)
print("kernel %.2f ms, %.2f us per step" % (dev.kernel_ms_, 1e3 * dev.kernel_ms_ / (104 * 8)))
