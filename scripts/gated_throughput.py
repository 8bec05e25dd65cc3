"""Throughput of the gated (tree-gate) SA path at the bench lattice with a synthetic depth-8 tree on the 7x7 window."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "structure-optimization_b200"))
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


d, n = 96, int(sys.argv[1]) if len(sys.argv) > 1 else 16384
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 8
lat = engine.Lattice(d, "NH3")
off = engine.local_offsets(3)
W1, b1, w2, b2 = glorot_mlp(off, 20, 0)
sweep = d * d
temps = exp_schedule(0.05, 5 * sweep)
# gate off (window kernel) / gated window kernel on the int32 and on the packed state / cached-gate kernel (variant 6) /
# row-per-lane kernel (variant 1)
cases = ((False, 0, False), (True, 0, False), (True, 0, True), (True, 6, False), (True, 1, False))
if len(sys.argv) > 3:
    cases = [cases[int(i)] for i in sys.argv[3].split(",")]
for gated, variant, compact in cases:
    tree = random_tree(49, depth, np.random.default_rng(0)) if gated else None
    tab = engine.SurrogateTables(lat, off, W1, b1, w2, b2, tree=tree, compact=compact)
    ch = engine.Chains(lat, n)
    ch.randomize(1, 0.5)
    ch.attach(tab)
    for i in range(3):          # sweeps 0, 1, 2 of a 5-sweep exponential schedule: acceptance falls as in a real run
        st = ch.anneal(temps[i * sweep:(i + 1) * sweep], seed=2, step0=i * sweep, variant=variant)
        print("depth %d gated=%s variant=%d compact=%d %s sweep %d: %.3e flip-evals/s (kernel %.1f ms, accept %.3f)" % (
            depth, gated, variant, compact, st["kernel"], i, st["flips"] / st["kernel_ms"] * 1e3, st["kernel_ms"],
            st["accepted"] / st["flips"]), flush=True)
    ch.close()
