"""Soak test of the pipelined SA kernel at full bench size with HIGH acceptance (hazard paths exercised ~1e7 times):
after many sweeps the tracked objective sums and the first-layer state of every chain must equal a fresh recompute
from the final bits, and a sample of chains must equal the C oracle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "structure-optimization_b200"))
import numpy as np
from so_b200 import engine
from bench import glorot_mlp
d, n, sweeps = 96, int(sys.argv[1]) if len(sys.argv) > 1 else 65536, int(sys.argv[2]) if len(sys.argv) > 2 else 6
N = d * d
lat = engine.Lattice(d, "NH3")
off = engine.local_offsets(3)
W1, b1, w2, b2 = glorot_mlp(off, 20, 0)
tab = engine.SurrogateTables(lat, off, W1, b1, w2, b2)
ch = engine.Chains(lat, n)
ch.randomize(7, 0.5)
x0 = ch.get_occupancy(0, 64)
ch.attach(tab)
flips = acc = 0
t0 = time.time()
for s in range(sweeps):
    T = 1.0 if s % 2 == 0 else 0.05                      # alternate hot (a ~ 0.45) and cold sweeps
    st = ch.anneal(np.full(N, T), step0=s * N, seed=5)
    flips += st["flips"]; acc += st["accepted"]
print("ran %d flip-evaluations, accept fraction %.3f, %.1f s" % (flips, acc / flips, time.time() - t0))
of, hi, lo, na = ch.objective(sums=True)
sample = [0, 1, n // 2, n - 1]
hq_before = [ch.preact(c) for c in sample]
fresh = ch.score()
of2, hi2, lo2, na2 = ch.objective(sums=True)
assert (hi == hi2).all() and (lo == lo2).all() and (of == fresh).all(), "tracked sums differ from a fresh recompute"
for c, h in zip(sample, hq_before):
    assert (ch.preact(c) == h).all(), "first-layer state of chain %d differs from a fresh recompute" % c
print("state invariant holds for all %d chains" % n)
# oracle parity on the first chains over the whole run
from oracle import c_oracle, surrogate as OS, lattice as OL, philox
p = OS.SurrogateParams(OL.local_offsets(3), W1, b1, w2, b2)
q = OS.quantize(p)
k = 8
steps = sweeps * N
prop, u = philox.draws(5, np.arange(k)[:, None], np.arange(steps)[None, :], N)
temps = np.concatenate([np.full(N, 1.0 if s % 2 == 0 else 0.05) for s in range(sweeps)])
o = c_oracle.anneal_fixed_many(p, q, d, x0[:k], temps, prop, u)
x1 = ch.get_occupancy(0, k)
assert (x1 == o["x"]).all() and (of[:k] == o["OF"]).all(), "chains differ from the oracle"
print("first %d chains equal the oracle after %d steps" % (k, steps))
