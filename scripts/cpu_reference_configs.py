"""CPU baselines B1-B5 of BASELINE.md section 3, measured with the oracle's restatement of the reference (CPU only; the
reference itself cannot be imported: Python 2 + ASE + zacros_wrapper).  Prints one JSON line per baseline."""
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(k, "1")
import numpy as np
from oracle import lattice as L, surrogate as S, sa as SA, philox


def b1_chain(seed, steps=14400, d=12):
    import helpers as Hp
    p = Hp.golden_surrogate(gated=True)

    class Sur(object):
        def eval_rate(self, M):
            return S.eval_rate_fp64(p, M)
    rng = np.random.default_rng(seed)
    x0 = np.zeros(d * d, dtype=np.uint8)
    x0[rng.permutation(d * d)[:d * d // 2]] = 1                       # randomize(coverage=0.5): exactly N/2 ones
    prop, u = philox.draws(seed, seed, np.arange(steps), d * d)
    t0 = time.perf_counter()
    SA.optimize_faithful(x0, d, Sur(), p.offsets, SA.schedule("exp", 0.05, steps), prop, u)
    return steps / (time.perf_counter() - t0)


def main():
    cores = os.cpu_count() or 1
    out = []
    r = b1_chain(0)
    out.append(dict(id="B1", what="faithful reference algorithm, 1 chain, d=12, trained 144->20->1 MLP + tree gate, "
                                  "n_cycles=100 (14,400 steps)", cores=1, value=r, unit="flip-evals/s"))
    with mp.get_context("spawn").Pool(cores) as pool:
        t0 = time.perf_counter()
        pool.map(b1_chain, range(cores))
        wall = time.perf_counter() - t0
    out.append(dict(id="B2", what="B1 x one independent chain per host core", cores=cores, value=cores * 14400 / wall,
                    unit="flip-evals/s"))
    # B3: the faithful algorithm at d=96 with the full translation descriptor (K = N = 9216): 680 MB matrix per chain
    d = 96
    p = S.SurrogateParams.random_init(L.full_offsets(d), H=20, seed=0)

    class Sur(object):
        def eval_rate(self, M):
            return S.eval_rate_fp64(p, M)
    x0 = (np.random.default_rng(0).random(d * d) < 0.5).astype(np.uint8)
    steps = 20
    prop, u = philox.draws(1, 0, np.arange(steps), d * d)
    syms = L.descriptor_rows(x0, p.offsets, d)
    t0 = time.perf_counter()
    SA.optimize_faithful(x0, d, Sur(), p.offsets, SA.schedule("exp", 0.05, steps), prop, u, syms=syms)
    out.append(dict(id="B3", what="faithful reference algorithm at d=96, full descriptor K=N=9216 (680 MB translation "
                                  "matrix copied per step), %d steps measured" % steps, cores=1,
                    value=steps / (time.perf_counter() - t0), unit="flip-evals/s"))
    # B4: batch scoring, d=12, 4096 structures
    import helpers as Hp
    pg = Hp.golden_surrogate(gated=True)
    xs = Hp.random_structures(4096, 144, seed=2)
    t0 = time.perf_counter()
    for x in xs:
        S.eval_rate_fp64(pg, L.generate_all_translations(x, 12))
    out.append(dict(id="B4", what="batch scoring (translation matrix + tree + MLP) of 4,096 structures, d=12", cores=1,
                    value=4096 / (time.perf_counter() - t0), unit="structures/s"))
    # B5: optimised CPU path (NOT the reference): C + OpenMP delta algorithm, local 7x7 descriptor at d=96
    os.environ["OMP_NUM_THREADS"] = str(cores)
    from oracle import c_oracle
    pl = S.SurrogateParams.random_init(L.local_offsets(3), H=20, seed=0)
    ql = S.quantize(pl)
    nc, ns = 4 * cores, 20000
    x0 = (np.random.default_rng(0).random((nc, d * d)) < 0.5).astype(np.uint8)
    prop, u = philox.draws(1, np.arange(nc)[:, None], np.arange(ns)[None, :], d * d)
    t0 = time.perf_counter()
    c_oracle.anneal_fixed_many(pl, ql, d, x0, SA.schedule("exp", 0.05, ns), prop, u)
    out.append(dict(id="B5", what="optimised CPU (not the reference): delta algorithm in C + OpenMP, d=96, K=49", cores=cores,
                    value=nc * ns / (time.perf_counter() - t0), unit="flip-evals/s"))
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
