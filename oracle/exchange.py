"""
ORACLE (test infrastructure, NOT product code) -- parallel-tempering swap round and best-structure pick.
Neither exists in the reference (it only scatters job lists over MPI ranks,
drivers/generate_init_structures.py:154-160); the semantics are those of SURVEY 8(e), restated here so that the
device kernels (exchange_kernel / best_kernel in structure-optimization_b200/csrc/so_sa.cu) have a checker.
"""
import math
import numpy as np
from . import philox


def gathered_position(g, world, n_local):
    return (g % world) * n_local + g // world


def exchange_round(of_all, tscale_all, world, n_local, ladder, parity, seed, rnd, t_ref, ladders=None):
    """Returns the new tscale array (all-gather layout [world][n_local]).  Global chain g = i*world + rank; ladders
    are `ladder` consecutive global ids; rungs ranked by temperature (hottest first, ties by global id); adjacent
    rungs (r, r+1), r = parity (mod 2), swap temperatures with probability min(1, exp[(OF_a-OF_b)(1/T_b-1/T_a)])."""
    of_all = np.asarray(of_all, dtype=np.float64)
    ts = np.asarray(tscale_all, dtype=np.float64)
    new = ts.copy()
    n_global = world * n_local
    seed = int(seed)
    for lad in (range(n_global // ladder) if ladders is None else [int(v) for v in ladders]):   # `ladders`: a sample
        pos = [gathered_position(lad * ladder + i, world, n_local) for i in range(ladder)]
        order = sorted(range(ladder), key=lambda i: (-ts[pos[i]], i))
        for r in range(parity, ladder - 1, 2):
            a, b = pos[order[r]], pos[order[r + 1]]
            Ta, Tb = ts[a] * t_ref, ts[b] * t_ref
            arg = (of_all[a] - of_all[b]) * (1.0 / Tb - 1.0 / Ta)
            out = philox.philox4x32_10((np.array([lad & 0xffffffff]), np.array([lad >> 32]), np.array([rnd & 0xffffffff]),
                                        np.array([r])),
                                       (np.array([(seed & 0xffffffff) ^ 0x5bd1e995]), np.array([seed >> 32])))
            u = float(int(out[0][0]) >> 8) * 2.0 ** -24
            if arg >= 0.0 or math.exp(arg) > u:
                new[a], new[b] = ts[b], ts[a]
    return new
