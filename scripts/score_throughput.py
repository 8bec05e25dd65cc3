"""State-build (first layer + objective from bits) throughput at the bench shape; SO_SCORE_TC=0 selects the CUDA-core kernel."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "structure-optimization_b200"))
import numpy as np
import torch
from so_b200 import engine
from bench import glorot_mlp
d, n = 96, int(sys.argv[1]) if len(sys.argv) > 1 else 8192
lat = engine.Lattice(d, "NH3")
off = engine.local_offsets(3)
W1, b1, w2, b2 = glorot_mlp(off, 20, 0)
tab = engine.SurrogateTables(lat, off, W1, b1, w2, b2)
ch = engine.Chains(lat, n)
ch.randomize(1, 0.5)
ch.attach(tab)
res = {}
for mode in ("1", "0", "1", "0"):
    os.environ["SO_SCORE_TC"] = mode
    torch.cuda.synchronize(); t0 = time.perf_counter()
    of = ch.score()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    res.setdefault(mode, []).append((dt, of.copy(), ch.preact(n // 2).copy()))
    print("SO_SCORE_TC=%s: %.2f ms for %d chains (%.3e site-rows/s)" % (mode, dt * 1e3, n, n * d * d / dt))
assert (res["1"][-1][1] == res["0"][-1][1]).all() and (res["1"][-1][2] == res["0"][-1][2]).all()
print("tensor-core and CUDA-core state builds agree bit for bit")
