"""
oracle/train.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement of the regressor training the reference runs (OML/train_surrogate.py:141-152):
    MLPRegressor(activation='relu', learning_rate_init=0.0001, alpha=0.1, hidden_layer_sizes=(20,),
                 max_iter=10000, tol=0.00001).fit(X_reg, Y)
The arithmetic lives in scikit-learn (not vendored in the reference, un-pinned: README.rst:9); restated here from
its published algorithm -- `BaseMultilayerPerceptron._fit_stochastic` / `_backprop` / `_init_coef`,
`AdamOptimizer._get_updates`, `sklearn.utils.shuffle` -- and PINNED against the installed scikit-learn itself
(tests/test_oracle_train.py: same random_state => same loss curve and weights).

  init      Glorot uniform, bound sqrt(6 / (fan_in + fan_out)); draws in the order W1, b1, w2, b2
  epoch     sample_idx <- sample_idx[perm], perm = RandomState.shuffle(arange(n)); batches of min(200, n) in order
  batch     a = relu(X W1 + b1); p = a w2 + b2; delta2 = p - y
            loss = mean(delta2^2)/2 + alpha/2 (|W1|^2 + |w2|^2) / B
            grad w2 = (a^T delta2 + alpha w2)/B, grad b2 = mean(delta2)
            delta1 = (delta2 w2^T) * [a > 0];  grad W1 = (X^T delta1 + alpha W1)/B, grad b1 = mean(delta1)
  Adam      t += 1; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; lr_t = lr sqrt(1-b2^t)/(1-b1^t);
            p -= lr_t m / (sqrt(v) + eps)          (one t for all parameters)
  stop      loss_epoch = sum(batch_loss * B)/n; if loss > best - tol: count += 1 else count = 0; best = min(best, loss);
            stop when count > n_iter_no_change (10) or after max_iter epochs
"""
import numpy as np


def init_params(rng, K, H):
    """sklearn _init_coef order (layer 0 coef, layer 0 intercept, layer 1 coef, layer 1 intercept)."""
    b = np.sqrt(6.0 / (K + H))
    W1 = rng.uniform(-b, b, (K, H))
    b1 = rng.uniform(-b, b, H)
    b = np.sqrt(6.0 / (H + 1))
    w2 = rng.uniform(-b, b, (H, 1))
    b2 = rng.uniform(-b, b, 1)
    return W1, b1, w2.reshape(-1), float(b2[0])


def epoch_orders(rng, n, n_epochs, sample_idx=None):
    """Sample order of the next n_epochs epochs (sklearn.utils.shuffle applied cumulatively)."""
    if sample_idx is None:
        sample_idx = np.arange(n)
    out = np.empty((n_epochs, n), dtype=np.int32)
    for e in range(n_epochs):
        perm = np.arange(n)
        rng.shuffle(perm)
        sample_idx = sample_idx[perm]
        out[e] = sample_idx
    return out, sample_idx


def fit_adam(X, y, random_state=0, H=20, lr=1e-4, alpha=0.1, max_iter=200, tol=1e-5, beta1=0.9, beta2=0.999,
             eps=1e-8, n_iter_no_change=10, batch_size=None, init=None, orders=None):
    """Returns dict(W1, b1, w2, b2, loss_curve, n_iter).  `init` / `orders` override the RandomState draws."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    n, K = X.shape
    rng = np.random.RandomState(random_state) if not isinstance(random_state, np.random.RandomState) else random_state
    W1, b1, w2, b2 = init_params(rng, K, H) if init is None else [np.array(v, dtype=np.float64) for v in init]
    b2 = float(b2)
    B = min(200, n) if batch_size is None else int(np.clip(batch_size, 1, n))
    params = [W1, w2, b1, np.array([b2])]                 # sklearn order: coefs_ + intercepts_
    ms = [np.zeros_like(p) for p in params]
    vs = [np.zeros_like(p) for p in params]
    t = 0
    best, count, curve = np.inf, 0, []
    sample_idx = np.arange(n)
    for it in range(max_iter):
        if orders is None:
            perm = np.arange(n)
            rng.shuffle(perm)
            sample_idx = sample_idx[perm]
        else:
            sample_idx = np.asarray(orders[it])
        acc = 0.0
        for s0 in range(0, n, B):
            idx = sample_idx[s0:s0 + B]
            nb = len(idx)
            Xb, yb = X[idx], y[idx]
            W1, w2, b1, b2a = params
            a = np.maximum(Xb @ W1 + b1, 0.0)
            d2 = a @ w2 + b2a[0] - yb
            loss = (d2 ** 2).mean() / 2 + 0.5 * alpha * (np.dot(W1.ravel(), W1.ravel()) + np.dot(w2, w2)) / nb
            g_w2 = (a.T @ d2 + alpha * w2) / nb
            g_b2 = np.array([d2.sum() / nb])
            d1 = np.outer(d2, w2)
            d1[a == 0] = 0.0
            g_W1 = (Xb.T @ d1 + alpha * W1) / nb
            g_b1 = d1.sum(axis=0) / nb
            acc += loss * nb
            t += 1
            lr_t = lr * np.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
            for i, g in enumerate((g_W1, g_w2, g_b1, g_b2)):
                ms[i] = beta1 * ms[i] + (1 - beta1) * g
                vs[i] = beta2 * vs[i] + (1 - beta2) * (g ** 2)
                params[i] = params[i] + (-lr_t * ms[i] / (np.sqrt(vs[i]) + eps))
        cur = acc / n
        curve.append(cur)
        count = count + 1 if cur > best - tol else 0
        if cur < best:
            best = cur
        if count > n_iter_no_change:
            break
    W1, w2, b1, b2a = params
    return dict(W1=W1, b1=b1, w2=w2, b2=float(b2a[0]), loss_curve=np.array(curve), n_iter=len(curve))
