"""Soak of the gated SA kernels at scale, alternating hot and cold sweeps:
(a) d=12 full descriptor, golden MLP + tree, 65,536 chains, shared-memory-resident kernel (gate bits + path index);
(b) d=96 7x7 window, synthetic depth-8 tree, 16,384 / 32,768 chains, gated window kernel on the int32 / packed state.
After the run every chain's tracked sums / active count must equal a fresh recompute from its final bits, and the
first chains must equal the C oracle step for step (Philox streams)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "structure-optimization_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from so_b200 import engine
import helpers as Hp
from oracle import c_oracle, surrogate as OS, lattice as OL, philox
from test_gpu_parity import _random_tree


def soak(name, d, p, n, sweeps, steps_per_sweep, compact=False):
    N = d * d
    lat = engine.Lattice(d, "LSC")
    tab = engine.SurrogateTables(lat, p.offsets, p.W1, p.b1, p.w2, p.b2, p.y_mean, p.y_scale, p.tree, compact=compact)
    ch = engine.Chains(lat, n)
    ch.randomize(7, 0.5)
    x0 = ch.get_occupancy(0, 8)
    ch.attach(tab)
    flips = acc = 0
    t0 = time.time()
    temps_all = []
    for s in range(sweeps):
        T = 2.0 if s % 2 == 0 else 0.05
        temps = np.full(steps_per_sweep, T)
        temps_all.append(temps)
        st = ch.anneal(temps, step0=s * steps_per_sweep, seed=5)
        flips += st["flips"]; acc += st["accepted"]
    print("%s (%s): %d flip-evaluations, accept fraction %.3f, %.1f s" % (name, st["kernel"], flips, acc / flips, time.time() - t0), flush=True)
    of, hi, lo, na = ch.objective(sums=True)
    fresh = ch.score()
    of2, hi2, lo2, na2 = ch.objective(sums=True)
    assert (hi == hi2).all() and (lo == lo2).all() and (na == na2).all() and (of == fresh).all(), "tracked state differs from a fresh recompute"
    print("  state invariant holds for all %d chains (mean active rows %.1f of %d)" % (n, na.mean(), N))
    q = OS.quantize(p, bits="packed24") if compact else OS.quantize(p)
    k = 8
    steps = sweeps * steps_per_sweep
    prop, u = philox.draws(5, np.arange(k)[:, None], np.arange(steps)[None, :], N)
    o = c_oracle.anneal_fixed_many(p, q, d, x0[:k], np.concatenate(temps_all), prop, u)
    x1 = ch.get_occupancy(0, k)
    assert (x1 == o["x"]).all() and (of[:k] == o["OF"]).all(), "chains differ from the oracle"
    print("  first %d chains equal the C oracle after %d steps" % (k, steps), flush=True)
    ch.close()


soak("resident gated d=12", 12, Hp.golden_surrogate(gated=True), 65536, 6, 2880)
p = OS.SurrogateParams.random_init(OL.local_offsets(3), H=20, seed=1)
p.tree = _random_tree(49, 8, seed=2)
soak("gated window d=96, int32 state", 96, p, 16384, 4, 9216)
soak("gated window d=96, packed state", 96, p, 32768, 4, 9216, compact=True)
