"""
Generates the committed golden fixtures.  Run in the build container only (reads /root/reference for the one
fixture the reference ships; everything else comes from the oracle):

    python tests/golden/make_golden.py

  lattice_input_fixture.json   numbers of zacros_input/lattice_input.dat (cell, 14 sites, 2 neighbour pairs)
  surrogate_d12.npz            reference-recipe surrogate (tree + 144->20->1 MLP + Y scaler) trained with the
                               installed scikit-learn on 48 synthetic island structures labelled by the LSC
                               analytical model (BASELINE config 1), stored as plain arrays
  sa_d12_golden.npz            oracle SA trajectories (faithful fp64 + fixed-point) for seeded inputs
  types_golden.npz             LSC / NH3 site types of seeded structures (oracle closed forms == NetworkX literal)
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
from oracle import lattice as L, site_types as ST, surrogate as S, sa as SA, philox, lsc_analytical as LA


def lattice_fixture():
    path = "/root/reference/zacros_input/lattice_input.dat"
    lines = [l.strip() for l in open(path)]
    cell, coords, pairs = [], [], []
    i = lines.index("cell_vectors       # in row format (Angstroms)".strip())
    cell = [[float(v) for v in lines[i + 1].split()], [float(v) for v in lines[i + 2].split()]]
    i = [k for k, l in enumerate(lines) if l.startswith("site_coordinates")][0]
    for l in lines[i + 1:]:
        if not l:
            break
        coords.append([float(v) for v in l.split()])
    i = [k for k, l in enumerate(lines) if l.startswith("neighboring_structure")][0]
    for l in lines[i + 1:]:
        if l.startswith("end_"):
            break
        a, b = l.split()[0].split("-")
        pairs.append([int(a), int(b)])
    types = [int(v) for v in [l for l in lines if l.startswith("site_types")][0].split()[1:]]
    out = dict(source="zacros_input/lattice_input.dat", cell=cell, site_coordinates=coords, neighbor_pairs=pairs,
               site_types=types)
    json.dump(out, open(os.path.join(HERE, "lattice_input_fixture.json"), "w"), indent=1)
    print("lattice fixture:", len(coords), "sites")


def surrogate_fixture():
    d, N, n_strucs = 12, 144, 48
    occs, rates = [], []
    for i in range(n_strucs):
        x = LA.make_random_structure(i, n_strucs, d)
        occs.append(x)
        rates.append(LA.compute_site_rates_lsc(x, d))
    occs, rates = np.array(occs), np.array(rates)
    all_syms = np.vstack([L.generate_all_translations_and_rotations(x, d) for x in occs])
    sur = S.train(all_syms, rates, seed=0, max_iter=10000)
    p = S.SurrogateParams.from_sklearn(sur.predictor, sur.Y_scaler, sur.classifier, L.full_offsets(d))
    np.savez_compressed(os.path.join(HERE, "surrogate_d12.npz"), W1=p.W1, b1=p.b1, w2=p.w2, b2=p.b2, y_mean=p.y_mean,
                        y_scale=p.y_scale, t_feature=p.tree["feature"], t_thr=p.tree["thr"], t_left=p.tree["left"],
                        t_right=p.tree["right"], t_active=p.tree["active"], train_occs=occs, train_rates=rates)
    # sklearn cross-check of the extracted arrays on the training structures
    for x in occs[:6]:
        T = L.generate_all_translations(x, d)
        a, b = sur.eval_rate(T), S.eval_rate_fp64(p, T)
        assert abs(a - b) <= 1e-9 * max(1.0, abs(a)), (a, b)
    print("surrogate: tree nodes", len(p.tree["feature"]), "iters", sur.predictor.n_iter_, "max site rate", rates.max())
    return p


def load_surrogate(gated=True):
    z = np.load(os.path.join(HERE, "surrogate_d12.npz"))
    tree = dict(feature=z["t_feature"], thr=z["t_thr"], left=z["t_left"], right=z["t_right"], active=z["t_active"])
    return S.SurrogateParams(L.full_offsets(12), z["W1"], z["b1"], z["w2"], z["b2"], float(z["y_mean"]),
                             float(z["y_scale"]), tree if gated else None)


def sa_fixture(p):
    d, N = 12, 144
    steps, n_chains = 1440, 4
    rng = np.random.Generator(np.random.Philox(key=1234))
    x0 = (rng.random((n_chains, N)) < 0.5).astype(np.uint8)
    prop = rng.integers(0, N, size=(n_chains, steps)).astype(np.int32)
    u = rng.random((n_chains, steps), dtype=np.float32)
    temps = SA.schedule("exp", 0.05, steps)
    out = dict(x0=x0, prop=prop, u=u, temps=temps)
    for name, pp in (("gated", p), ("open", S.SurrogateParams(p.offsets, p.W1, p.b1, p.w2, p.b2, p.y_mean, p.y_scale, None))):
        q = S.quantize(pp)

        class Sur:
            def eval_rate(self, M, pp=pp):
                return S.eval_rate_fp64(pp, M)
        fa = [SA.optimize_faithful(x0[c], d, Sur(), pp.offsets, temps, prop[c], u[c]) for c in range(n_chains)]
        fx = [SA.anneal_fixed(pp, q, d, x0[c], temps, prop[c], u[c]) for c in range(n_chains)]
        out[name + "_faithful_x"] = np.array([r["x"] for r in fa])
        out[name + "_faithful_accept"] = np.array([r["accept"] for r in fa])
        out[name + "_faithful_OF"] = np.array([r["OF"] for r in fa])
        out[name + "_faithful_traj"] = np.array([r["traj"] for r in fa])
        out[name + "_fixed_x"] = np.array([r["x"] for r in fx])
        out[name + "_fixed_accept"] = np.array([r["accept"] for r in fx])
        out[name + "_fixed_OF"] = np.array([r["OF"] for r in fx])
        out[name + "_fixed_hi"] = np.array([r["hi"] for r in fx], dtype=np.int64)
        out[name + "_fixed_lo"] = np.array([r["lo"] for r in fx], dtype=np.int64)
        agree = [(a["accept"] == b["accept"]).all() for a, b in zip(fa, fx)]
        print("SA", name, "accept rate %.3f" % np.mean([r["accept"].mean() for r in fx]), "faithful==fixed:", agree)
    np.savez_compressed(os.path.join(HERE, "sa_d12_golden.npz"), **out)


def types_fixture():
    from nx_literal import lsc_types_nx, nh3_types_nx
    rng = np.random.default_rng(5)
    out = {}
    for d in (6, 8):
        xs = np.array([(rng.random(d * d) < p).astype(np.uint8) for p in (0.15, 0.4, 0.6, 0.85, 0.97)])
        lsc = np.array([lsc_types_nx(x, d)[3 * d * d:4 * d * d] for x in xs])
        nh3 = np.array([nh3_types_nx(x, d) for x in xs])
        assert (lsc == ST.lsc_types(xs, d)).all() and (nh3 == ST.nh3_types(xs, d)).all()
        out["x_d%d" % d], out["lsc_d%d" % d], out["nh3_d%d" % d] = xs, lsc, nh3
    np.savez_compressed(os.path.join(HERE, "types_golden.npz"), **out)
    print("types fixture ok")


if __name__ == "__main__":
    lattice_fixture()
    types_fixture()
    p = surrogate_fixture() if "--retrain" in sys.argv or not os.path.exists(os.path.join(HERE, "surrogate_d12.npz")) \
        else load_surrogate()
    sa_fixture(p)
