"""d=96 gated local-window SA (cached-gate kernel, global state): two warm sweeps, then a short launch for ncu
(--launch-skip 2 on the kernel regex)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "structure-optimization_b200")); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import numpy as np
from so_b200 import engine
from bench import glorot_mlp, exp_schedule


def random_tree(K, depth, rng):
    feature, thr, left, right, active = [], [], [], [], []

    def node(level):
        i = len(feature)
        feature.append(-2); thr.append(0.0); left.append(-1); right.append(-1); active.append(int(rng.random() < 0.6))
        if level < depth:
            feature[i] = int(rng.integers(0, K)); thr[i] = 0.5
            left[i] = node(level + 1); right[i] = node(level + 1)
        return i
    node(0)
    return dict(feature=np.array(feature, np.int32), thr=np.array(thr), left=np.array(left, np.int32),
                right=np.array(right, np.int32), active=np.array(active, np.uint8))


d, n = 96, 16384
lat = engine.Lattice(d, "NH3")
off = engine.local_offsets(3)
W1, b1, w2, b2 = glorot_mlp(off, 20, 0)
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 8
tab = engine.SurrogateTables(lat, off, W1, b1, w2, b2, tree=random_tree(49, depth, np.random.default_rng(0)))
ch = engine.Chains(lat, n); ch.randomize(1, 0.5); ch.attach(tab)
sweep = d * d
temps = exp_schedule(0.05, 5 * sweep)
for i in range(2):
    ch.anneal(temps[i * sweep:(i + 1) * sweep], seed=2, step0=i * sweep)
st = ch.anneal(temps[2 * sweep:2 * sweep + 1024], seed=2, step0=2 * sweep)
print("%.3e flip-evals/s accept %.3f" % (st["flips"] / st["kernel_ms"] * 1e3, st["accepted"] / st["flips"]))
