"""One launch of the resident SA kernel (d=12 full descriptor, gated unless 'open' is given) for ncu."""
import sys; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/structure-optimization_b200'); sys.path.insert(0, '/root/repo/tests')
from so_b200 import engine
import helpers as Hp
from oracle import sa as SA
gated = "open" not in sys.argv
n = 16384
lat = engine.Lattice(12, "LSC"); tab = Hp.tables_from_params(lat, Hp.golden_surrogate(gated=gated))
ch = engine.Chains(lat, n); ch.randomize(1, 0.5); ch.attach(tab)
temps = SA.schedule("exp", 0.05, 14400)[:300]
st = ch.anneal(temps, seed=3)
print("gated=%s %.3e flips/s" % (gated, st["flips"] / st["kernel_ms"] * 1e3))
