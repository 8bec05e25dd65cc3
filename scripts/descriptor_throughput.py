"""Device time of the descriptor kernels at the bench lattice (96x96): histograms only, and with the full per-node type
arrays; plus the per-flip delta kernel.  Times come from CUDA events inside a torch-free loop via cudaEvent in ctypes?
No: the C ABI is host-synchronous, so wall time around the call minus the D2H copy is reported (copy sizes printed)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "structure-optimization_b200"))
import numpy as np
from so_b200 import engine
d, n = 96, int(sys.argv[1]) if len(sys.argv) > 1 else 16384
N = d * d
lat = engine.Lattice(d, "NH3")
ch = engine.Chains(lat, n)
ch.randomize(1, 0.5)
for label, kw in (("histograms only", dict(lsc_hist=True, nh3_hist=True, nh3_exp=True)),
                  ("c6 + lsc types", dict(c6=True, lsc_type=True)),
                  ("all outputs", dict(c6=True, lsc_type=True, nh3_type=True, lsc_hist=True, nh3_hist=True, nh3_exp=True))):
    ch.descriptors(chain0=0, n=min(n, 1024), **kw)
    t0 = time.perf_counter(); out = ch.descriptors(**kw); dt = time.perf_counter() - t0
    nbytes = sum(v.nbytes for v in out.values())
    print("%-18s %8.2f ms wall for %d chains (%.1f MB out, %.1f MB bit-planes in)" % (label, dt * 1e3, n, nbytes / 1e6, n * lat.W * 4 / 1e6))
sites = np.random.default_rng(0).integers(0, N, n).astype(np.int32)
ch.flip_delta(sites[:1024])
t0 = time.perf_counter(); ch.flip_delta(sites); dt = time.perf_counter() - t0
print("flip_delta         %8.2f ms wall for %d (chain, site) pairs" % (dt * 1e3, n))
