"""Two launches of the resident SA kernel (d=12 full descriptor, gated unless 'open' is given); ncu captures the second
(steady state: annealed structures, low acceptance) with --launch-skip 1."""
import sys; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/structure-optimization_b200'); sys.path.insert(0, '/root/repo/tests')
from so_b200 import engine
import helpers as Hp
from oracle import sa as SA
gated = "open" not in sys.argv
n = 16384
lat = engine.Lattice(12, "LSC"); tab = Hp.tables_from_params(lat, Hp.golden_surrogate(gated=gated))
ch = engine.Chains(lat, n); ch.randomize(1, 0.5); ch.attach(tab)
temps = SA.schedule("exp", 0.05, 14400)
ch.anneal(temps[:7200], seed=3)
st = ch.anneal(temps[7200:7500], seed=3, step0=7200)
of, hi, lo, na = ch.objective(sums=True)
print("mean active rows per chain: %.1f of %d" % (na.mean(), lat.N))
print("gated=%s %.3e flips/s accept %.4f" % (gated, st["flips"] / st["kernel_ms"] * 1e3, st["accepted"] / st["flips"]))
