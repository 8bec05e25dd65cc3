"""Few-to-many chains on the 12x12 lattice: CTA-per-chain kernel (variant 5) against warp-per-chain (variant 4);
the break-even (about two chains per SM) is what so_team_ok() encodes."""
import sys; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/structure-optimization_b200'); sys.path.insert(0, '/root/repo/tests')
from so_b200 import engine
import helpers as Hp
from oracle import sa as SA
for gated in (True, False):
    lat = engine.Lattice(12, "LSC"); tab = Hp.tables_from_params(lat, Hp.golden_surrogate(gated=gated))
    temps = SA.schedule("exp", 0.05, 14400)
    for n in (64, 296, 600, 740, 1000):
        for v in (4, 5):
            ch = engine.Chains(lat, n); ch.randomize(1, 0.5); ch.attach(tab)
            ch.anneal(temps, seed=3, variant=v)
            st = ch.anneal(temps, seed=4, variant=v)
            print("gated=%s chains=%d variant=%d: %.3e flips/s kernel %.1f ms" % (gated, n, v, st["flips"]/st["kernel_ms"]*1e3, st["kernel_ms"]), flush=True)
            ch.close()
