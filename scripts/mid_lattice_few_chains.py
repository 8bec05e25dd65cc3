import sys; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/structure-optimization_b200'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from so_b200 import engine
import helpers as Hp
from oracle import sa as SA, surrogate as S, lattice as L
from test_gpu_parity import _random_tree
for d in (32, 48):
    for gated in (False, True):
        p = S.SurrogateParams.random_init(L.local_offsets(3), H=20, seed=3)
        if gated:
            p.tree = _random_tree(49, 7, seed=4)
        lat = engine.Lattice(d, "NH3"); tab = Hp.tables_from_params(lat, p)
        temps = SA.schedule("exp", 0.05, 4000)
        for n in (1, 100):
            res = []
            for v in (0, 3):
                ch = engine.Chains(lat, n); ch.randomize(1, 0.5); ch.attach(tab)
                ch.anneal(temps, seed=3, variant=v)
                st = ch.anneal(temps, seed=4, variant=v)
                res.append((ch.get_bits().copy(), ch.objective().copy()))
                print("d=%d gated=%s chains=%d variant=%d: %.2f us per step" % (d, gated, n, v, 1e3 * st["kernel_ms"] / len(temps)), flush=True)
                ch.close()
            assert (res[0][0] == res[1][0]).all() and (res[0][1] == res[1][1]).all(), "variants disagree"
