"""NH3_cat's default lattice (16x16, N = 256) with the full 256-column descriptor, gate on and off: which kernel wins
where the state (20 KB per chain + 8 KB path index) only just fits in shared memory."""
import sys; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/structure-optimization_b200'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from so_b200 import engine
import helpers as Hp
from oracle import sa as SA, surrogate as S, lattice as L
from test_gpu_parity import _random_tree
d = 16
for gated in (True, False):
    p = S.SurrogateParams.random_init(L.full_offsets(d), H=20, seed=3)
    if gated:
        p.tree = _random_tree(d * d, 10, seed=4, grow=0.7)
    lat = engine.Lattice(d, "NH3"); tab = Hp.tables_from_params(lat, p)
    temps = SA.schedule("exp", 0.05, 5 * d * d)
    for n in (1, 64, 4096, 32768):
        for v in ((0, 4, 5, 1) if gated else (0, 4, 5, 2)):
            if v == 5 and n > 4096:
                continue
            ch = engine.Chains(lat, n); ch.randomize(1, 0.5); ch.attach(tab)
            ch.anneal(temps, seed=3, variant=v)
            st = ch.anneal(temps, seed=4, variant=v)
            print("gated=%s chains=%5d variant=%d: %.3e flips/s kernel %.1f ms (%.2f us per step of one chain-slot)" % (
                gated, n, v, st["flips"] / st["kernel_ms"] * 1e3, st["kernel_ms"], 1e3 * st["kernel_ms"] / len(temps)), flush=True)
            ch.close()
