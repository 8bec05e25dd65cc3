"""Shared test helpers: golden loaders and oracle <-> product glue."""
import os
import numpy as np
from oracle import lattice as L, surrogate as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_surrogate(gated=True):
    z = np.load(os.path.join(GOLDEN, "surrogate_d12.npz"))
    tree = dict(feature=z["t_feature"], thr=z["t_thr"], left=z["t_left"], right=z["t_right"], active=z["t_active"])
    return S.SurrogateParams(L.full_offsets(12), z["W1"], z["b1"], z["w2"], z["b2"], float(z["y_mean"]),
                             float(z["y_scale"]), tree if gated else None)


def random_structures(n, N, seed, coverage=None):
    rng = np.random.default_rng(seed)
    p = rng.random((n, 1)) if coverage is None else np.full((n, 1), coverage)
    return (rng.random((n, N)) < p).astype(np.uint8)


def tables_from_params(lattice, p):
    from so_b200.engine import SurrogateTables
    return SurrogateTables(lattice, p.offsets, p.W1, p.b1, p.w2, p.b2, p.y_mean, p.y_scale, p.tree)


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
