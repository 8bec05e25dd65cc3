"""optimize(mode='analytical') on the device (so_anneal_analytical): Metropolis steps/s, each step = site typing + assembly +
solve of the LSC steady-state system, next to the oracle's literal loop (LAPACK solve per step) on one host core."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "structure-optimization_b200"))
import numpy as np
from so_b200 import engine
from oracle import sa as SA
for d, n, steps in ((12, 2048, 200), (16, 2048, 100), (24, 1024, 100), (96, 296, 20)):
    N = d * d
    rng = np.random.default_rng(d)
    xs = (rng.random((n, N)) < 0.85).astype(np.uint8)
    ch = engine.Chains(engine.Lattice(d, "LSC"), n)
    ch.set_occupancy(xs)
    temps = SA.schedule("exp", 0.05, steps)
    ch.anneal_analytical(temps[:4], seed=1)
    st = ch.anneal_analytical(temps, seed=2)
    gpu = st["flips"] / (st["kernel_ms"] * 1e-3)
    k = max(5, min(steps, 2000 // d))
    prop = rng.integers(0, N, k); u = rng.random(k, dtype=np.float32)
    t0 = time.perf_counter(); SA.optimize_analytical(xs[0], d, temps[:k], prop, u); cpu = k / (time.perf_counter() - t0)
    print("d=%3d (%5d sites, %s): %5d chains x %3d steps  %9.1f ms  -> %.3e steps/s on the GPU; oracle (LAPACK, 1 core) %.1f steps/s; x%.0f"
          % (d, N, "dense elimination in shared memory" if N <= 160 else "CG on the lattice stencil", n, steps,
             st["kernel_ms"], gpu, cpu, gpu / cpu), flush=True)
