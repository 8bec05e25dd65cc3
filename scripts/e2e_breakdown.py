"""Where the end-to-end step of bench.py spends its time beyond the sweep: set_bits (H2D + state build), anneal, get_bits, objective."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "structure-optimization_b200"))
import numpy as np, torch
import bench
from so_b200 import engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
d, N = 96, 96 * 96
lat = engine.Lattice(d, "NH3", device=0)
off = engine.local_offsets(3)
W1, b1, w2, b2 = bench.glorot_mlp(off, bench.H, 0)
tab = engine.SurrogateTables(lat, off, W1, b1, w2, b2, compact=True)
ch = engine.Chains(lat, n)
ch.randomize(seed=1, coverage=0.5)
ch.attach(tab)
temps = bench.exp_schedule(bench.T0, 6 * N)
for i in range(3):
    ch.anneal(temps[i * N:(i + 1) * N], step0=i * N, seed=7)
bits = torch.empty((n, lat.W), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
of = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
ch.get_bits(out=bits)
for rep in range(3):
    t = [time.perf_counter()]
    ch.set_bits(bits); torch.cuda.synchronize(); t.append(time.perf_counter())
    st = ch.anneal(temps[-N:], step0=(3 + rep) * N, seed=7); t.append(time.perf_counter())
    ch.get_bits(out=bits); t.append(time.perf_counter())
    ch.objective(out=of); t.append(time.perf_counter())
    ms = [1e3 * (b - a) for a, b in zip(t[:-1], t[1:])]
    print("set_bits %.2f ms  anneal %.2f ms (kernel %.2f)  get_bits %.2f ms  objective %.2f ms  total %.2f ms" % (
        ms[0], ms[1], st["kernel_ms"], ms[2], ms[3], sum(ms)), flush=True)
