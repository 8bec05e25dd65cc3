"""Equilibrium accept fraction of the bench workload (96x96, 7x7-window MLP, Glorot seed 0) as a function of a CONSTANT
temperature: two sweeps per temperature (one to equilibrate, one measured).  bench.py's accept_sweep temperatures come
from this table (profiles/r02_accept_calibration.txt)."""
import sys, os, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "structure-optimization_b200"))
import bench
from so_b200 import engine

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
d, N = 96, 96 * 96
lat = engine.Lattice(d, "NH3", device=0)
off = engine.local_offsets(3)
W1, b1, w2, b2 = bench.glorot_mlp(off, bench.H, 0)
tab = engine.SurrogateTables(lat, off, W1, b1, w2, b2)
ch = engine.Chains(lat, chains)
ch.randomize(seed=1234, coverage=0.5)
ch.attach(tab)
step0 = 0
temps = [float(v) for v in os.environ['TEMPS'].split(',')] if os.environ.get('TEMPS') else np.logspace(1, -2.5, 15)
for T in temps:
    for rep in range(2):
        st = ch.anneal(np.full(N, T), step0=step0, seed=7, chain_id0=0)
        step0 += N
    a = st["accepted"] / st["flips"]
    b = 4.0 * 49 * 20 * (1 + a) + 15 + 4 * a
    print("T=%9.5f accept=%.4f  %.1f ms  %.3e flip-evals/s  %.0f GB/s (algorithmic)" % (
        T, a, st["kernel_ms"], st["flips"] / st["kernel_ms"] * 1e3, st["flips"] * b / st["kernel_ms"] / 1e6), flush=True)
