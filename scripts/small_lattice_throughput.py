"""Small-lattice SA (the reference's own sizes): resident kernel (variant 0) against the global-memory kernels."""
import sys; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/structure-optimization_b200'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from so_b200 import engine
import helpers as Hp
from oracle import sa as SA, surrogate as S, lattice as L

def run(name, d, p, n_chains, steps, variants):
    lat = engine.Lattice(d, "LSC"); tab = Hp.tables_from_params(lat, p)
    temps = SA.schedule("exp", 0.05, steps)
    for v in variants:
        ch = engine.Chains(lat, n_chains); ch.randomize(1, 0.5); ch.attach(tab)
        ch.anneal(temps, seed=3, variant=v)
        st = ch.anneal(temps, seed=4, variant=v)
        print("%-28s d=%d chains=%6d variant=%d: %.3e flips/s  kernel %.1f ms accept %.3f" % (
            name, d, n_chains, v, st["flips"] / st["kernel_ms"] * 1e3, st["kernel_ms"], st["accepted"] / st["flips"]), flush=True)
        ch.close()

for n in (4096, 65536):
    steps = 14400 if n == 4096 else 1440
    run("full descriptor open", 12, Hp.golden_surrogate(gated=False), n, steps, (0, 2))
    run("full descriptor gated", 12, Hp.golden_surrogate(gated=True), n, steps, (0, 1))
    pl = S.SurrogateParams.random_init(L.local_offsets(3), H=20, seed=3)
    run("local 7x7 open", 12, pl, n, steps, (0, 3))
    run("local 7x7 open", 16, pl, n, steps, (0, 3))
