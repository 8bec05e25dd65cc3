"""Descriptor kernels alone, for ncu captures and device timing: python scripts/descriptor_probe.py d n_structures
  descriptors_bits_kernel   LSC / NH3 histograms of n structures (bit-sliced)
  descriptors_kernel        per-node outputs (c6, LSC types, 19 NH3 types)
  flip_delta_kernel         histogram deltas of one proposed flip per structure"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "structure-optimization_b200"))
import numpy as np
from so_b200 import engine
d = int(sys.argv[1]) if len(sys.argv) > 1 else 96
n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
N = d * d
lat = engine.Lattice(d, "NH3")
ch = engine.Chains(lat, n)
ch.randomize(1, 0.5)
sites = np.random.default_rng(0).integers(0, N, n).astype(np.int32)
for rep in range(3):
    t0 = time.perf_counter(); out = ch.descriptors(lsc_hist=True, nh3_hist=True, nh3_exp=True); t1 = time.perf_counter()
    dl, dn = ch.flip_delta(sites); t2 = time.perf_counter()
print("d=%d n=%d: histograms %.2f ms wall (%.1f MB bit-planes in, %.1f MB out); flip_delta %.2f ms wall" % (
    d, n, (t1 - t0) * 1e3, n * lat.W * 4 / 1e6, sum(v.nbytes for v in out.values()) / 1e6, (t2 - t1) * 1e3))
m = min(n, 4096)
ch.descriptors(chain0=0, n=m, c6=True, lsc_type=True, nh3_type=True)
