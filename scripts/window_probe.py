"""One or more sweeps of the headline kernel at a reduced chain count (for ncu captures): python scripts/window_probe.py
[chains] [sweeps] [T or 'cool']."""
import sys, os, ctypes
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "structure-optimization_b200"))
import bench
from so_b200 import engine

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
T = sys.argv[3] if len(sys.argv) > 3 else "cool"
d, N = 96, 96 * 96
lat = engine.Lattice(d, "NH3", device=0)
rt = ctypes.CDLL("libcudart.so.12")
val = ctypes.c_size_t()
rt.cudaDeviceGetLimit(ctypes.byref(val), 7)          # cudaLimitMaxL2FetchGranularity = 0x07
print("cudaLimitMaxL2FetchGranularity reads back as", val.value)
off = engine.local_offsets(3)
W1, b1, w2, b2 = bench.glorot_mlp(off, bench.H, 0)
tab = engine.SurrogateTables(lat, off, W1, b1, w2, b2, compact=os.environ.get('SO_COMPACT', '1') != '0')
ch = engine.Chains(lat, chains)
ch.randomize(seed=1234, coverage=0.5)
ch.attach(tab)
temps = bench.exp_schedule(bench.T0, sweeps * N) if T == "cool" else np.full(sweeps * N, float(T))
for s in range(sweeps):
    st = ch.anneal(temps[s * N:(s + 1) * N], step0=s * N, seed=7, chain_id0=0)
    a = st["accepted"] / st["flips"]
    print("sweep %d: %s accept=%.4f %.2f ms %.3e flip-evals/s" % (s, st["kernel"], a, st["kernel_ms"], st["flips"] / st["kernel_ms"] * 1e3))
