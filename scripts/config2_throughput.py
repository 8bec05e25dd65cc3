import sys, time; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/structure-optimization_b200'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from so_b200 import engine
import helpers as Hp
from oracle import sa as SA
for gated in (False, True):
    p = Hp.golden_surrogate(gated=gated)
    lat = engine.Lattice(12, "LSC"); tab = Hp.tables_from_params(lat, p)
    ch = engine.Chains(lat, 4096); ch.randomize(1, 0.5); ch.attach(tab)
    temps = SA.schedule("exp", 0.05, 14400)
    st = ch.anneal(temps, seed=3)
    st = ch.anneal(temps, seed=4)
    print("config2 gated=%s: %.3e flips/s  kernel %.1f ms accept %.3f" % (gated, st["flips"]/st["kernel_ms"]*1e3, st["kernel_ms"], st["accepted"]/st["flips"]))
