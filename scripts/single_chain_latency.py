"""Latency of the reference's own use case: ONE chain, d=12, trained 144->20->1 MLP + tree gate, n_cycles = 100 / 500
(drivers/Online_learn_main.py:186-187), through the drop-in optimize()."""
import os, sys, time, random, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "structure-optimization_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from so_b200 import LSC_cat, optimize
import helpers as Hp
from test_gpu_api import duck_surrogate
for gated in (True, False):
    sur = duck_surrogate(Hp.golden_surrogate(gated=gated))
    for n_cycles, n_chains, guard in ((100, 1, True), (500, 1, True), (500, 1, False), (500, 4096, False)):
        cat = LSC_cat(12)
        random.seed(0)
        cat.randomize(coverage=0.5)
        t0 = time.perf_counter()
        out = optimize(cat, surrogate=sur, n_cycles=n_cycles, T_0=0.05, n_chains=n_chains, seed=1, exact_guard=guard, return_stats=True)
        dt = time.perf_counter() - t0
        print("gated=%s n_cycles=%d n_chains=%d guard=%s: %.3f s wall, kernels %.1f ms, %.1f us/step/chain-launch, OF %.4f, guard fired %d"
              % (gated, n_cycles, n_chains, guard, dt, out["kernel_ms"], 1e3 * out["kernel_ms"] / (n_cycles * 144), out["OF"], out["guard_fired"]))
