"""d = 12, full 144-column descriptor, tree gate: the golden surrogate (33 % of the rows active) and bench.py's synthetic depth-8
tree (60 % active leaves), 65,536 chains x 1,440 steps, second launch timed.  SO_RES_ROWLANE=0/1 forces the difference pass."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "structure-optimization_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from so_b200 import engine
import helpers as Hp
import bench
from oracle import sa as SA

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
d, N = 12, 144
lat = engine.Lattice(d, "LSC")
cases = []
cases.append(("golden", Hp.tables_from_params(lat, Hp.golden_surrogate(gated=True))))
rng = np.random.default_rng(0)
t = dict(feature=[], thr=[], left=[], right=[], active=[])
def node(level):
    i = len(t["feature"])
    for k, v in (("feature", -2), ("thr", 0.0), ("left", -1), ("right", -1), ("active", int(rng.random() < 0.6))):
        t[k].append(v)
    if level < 8:
        t["feature"][i], t["thr"][i] = int(rng.integers(0, N)), 0.5
        t["left"][i] = node(level + 1); t["right"][i] = node(level + 1)
    return i
node(0)
tree = dict(feature=np.array(t["feature"], np.int32), thr=np.array(t["thr"]), left=np.array(t["left"], np.int32),
            right=np.array(t["right"], np.int32), active=np.array(t["active"], np.uint8))
off = engine.full_offsets(d)
W1, b1, w2, b2 = bench.glorot_mlp(off, 20, 0)
cases.append(("synthetic depth 8", engine.SurrogateTables(lat, off, W1, b1, w2, b2, tree=tree)))
temps = SA.schedule("exp", 0.05, 2880)
for name, tab in cases:
    ch = engine.Chains(lat, n); ch.randomize(1, 0.5); ch.attach(tab)
    ch.anneal(temps[:1440], seed=3)
    st = ch.anneal(temps[1440:], seed=3, step0=1440)
    na = ch.objective(sums=True)[3]
    print("%-18s ROWLANE=%s %s: %.3e flip-evals/s (%.1f ms, accept %.4f, mean active rows %.1f)" % (
        name, os.environ.get("SO_RES_ROWLANE", "default"), st["kernel"], st["flips"] / st["kernel_ms"] * 1e3, st["kernel_ms"],
        st["accepted"] / st["flips"], na.mean()), flush=True)
    ch.close()
