"""
Generates tests/golden/ref_*.npz by EXECUTING THE REFERENCE'S OWN CODE (tests/ref_exec.py lists the mechanical
py2->py3 transformations and the shims; nothing here comes from oracle/).  Run in the build container only:

    python tests/golden/make_ref_golden.py            # ~3 min

  ref_dynamic_cat.npz    OML/dynamic_cat.py: translate, rotate, generate_all_translations(_and_rotations),
                         get_local_inds, index maps, randomize, flip_atom/rand_move/revert_last, geo_crossover,
                         calc_weights, dyn_cat_dist
  ref_types.npz          OML/LSC_cat.py + OML/NH3_cat.py graph_to_KMClattice(): site_type of all 5N graph nodes,
                         KMC_lat.site_type_inds, var_ind_kmc_sites, KMC neighbour list; compute_site_rates_lsc(),
                         eval_struc_rate_anal(); d = 6, 8, 12, 16
  ref_surrogate_sa.npz   drivers/generate_init_structures.py make_random_structure (48 structures), LSC analytical
                         labels, OML/train_surrogate.py partition_data_set / train_classifier / train_regressor,
                         eval_structure_rate scores, and OML/optimize_SA.py optimize() runs with injected
                         proposal / uniform sequences (accept vectors, objective per step, final occupancies,
                         recorded trajectory) -- gated and always-active classifier, cold and hot T_0, all four
                         cooling schedules, mode='analytical'

After writing, every array is compared with the oracle's restatement (tests/test_ref_golden.py does the same in CI).
"""
import os
import random
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import ref_exec


def ragged(lists, dtype=np.int32):
    ptr = np.zeros(len(lists) + 1, dtype=np.int64)
    ptr[1:] = np.cumsum([len(v) for v in lists])
    flat = np.array([e for v in lists for e in v], dtype=dtype)
    return ptr, flat


def dynamic_cat_fixture(R):
    out = {}
    for d in (6, 12):
        cat = R.LSC_cat(dimen=d)
        N = d * d
        rng = np.random.default_rng(100 + d)
        x = (rng.random(N) < 0.45).astype(int).tolist()
        cat.variable_occs = list(x)
        tag = "d%d_" % d
        out[tag + "x"] = np.array(x, dtype=np.uint8)
        shifts = [(0, 0), (1, 0), (0, 1), (-1, 2), (d + 3, -2 * d - 1), (5, 4)]
        out[tag + "shifts"] = np.array(shifts, dtype=np.int32)
        out[tag + "translate"] = np.array([cat.translate(a, b) for a, b in shifts])
        out[tag + "rotate"] = np.array([np.asarray(cat.rotate(k), dtype=float) for k in (0, 1, 2, 4)])
        out[tag + "local_inds"] = np.array([cat.get_local_inds(a, b) for a, b in shifts], dtype=np.int32)
        out[tag + "var2sym"] = np.array([cat.var_ind_to_sym_inds(v) for v in range(N)], dtype=np.int32)
        grid = [(a, b) for a in range(-d - 1, 2 * d + 2, 3) for b in range(-2 * d, d + 3, 2)]
        out[tag + "sym_in"] = np.array(grid, dtype=np.int32)
        out[tag + "sym2var"] = np.array([cat.sym_inds_to_var_ind(a, b) for a, b in grid], dtype=np.int32)
        if d == 6:
            out[tag + "all_trans"] = cat.generate_all_translations()
            out[tag + "all_syms"] = cat.generate_all_translations_and_rotations()
        else:
            out[tag + "all_trans_rows"] = cat.generate_all_translations()[[0, 1, 13, 77, 143]]
            out[tag + "all_syms_rows"] = cat.generate_all_translations_and_rotations()[[0, 1, 2, 40, 41, 431]]
        out[tag + "variable_atoms"] = np.array(list(cat.variable_atoms), dtype=np.int32)
        out[tag + "surface_area"] = np.float64(cat.surface_area)
        # randomize(): the global `random` stream, exactly as upstream consumes it
        R.set_random("dynamic_cat", random)
        random.seed(7)
        out[tag + "randomize_half"] = np.array(cat.randomize(coverage=0.5), dtype=np.uint8)
        out[tag + "randomize_none"] = np.array(cat.randomize(), dtype=np.uint8)
        out[tag + "randomize_third"] = np.array(cat.randomize(coverage=1.0 / 3), dtype=np.uint8)
        # rand_move / revert_last / flip_atom (atom indices)
        random.seed(11)
        cat.variable_occs = list(x)
        moved, states = [], []
        for k in range(12):
            cat.rand_move()
            moved.append(cat.atom_last_moved)
            if k % 3 == 2:
                cat.revert_last()
            states.append(list(cat.variable_occs))
        cat.flip_atom(3 * N + 5)
        states.append(list(cat.variable_occs))
        out[tag + "moved"] = np.array(moved, dtype=np.int32)
        out[tag + "move_states"] = np.array(states, dtype=np.uint8)
        # GA helpers
        random.seed(5)
        x2 = (rng.random(N) < 0.6).astype(int).tolist()
        c1, c2 = cat.geo_crossover(list(x), list(x2))
        c3, c4 = cat.geo_crossover(list(x), list(x2), pt1=2, pt2=3)
        out[tag + "xo_parent2"] = np.array(x2, dtype=np.uint8)
        out[tag + "xo_children"] = np.array([c1, c2, c3, c4], dtype=np.uint8)
        cat.variable_occs = list(x)
        w = cat.calc_weights()
        out[tag + "weights"] = w
        out[tag + "dist"] = np.float64(R.dyn_cat_dist(np.array(x, dtype=float), np.array(x2, dtype=float), w))
    base = R.dynamic_cat(dim1=6, dim2=6)            # the base class' own default: 3 fixed + 1 variable layer
    out["base_variable_atoms"] = np.array(list(base.variable_atoms), dtype=np.int32)
    out["base_occs"] = np.array(base.variable_occs, dtype=np.uint8)
    out["base_surface_area"] = np.float64(base.surface_area)
    np.savez_compressed(os.path.join(HERE, "ref_dynamic_cat.npz"), **out)
    print("ref_dynamic_cat.npz:", len(out), "arrays")


def structures_for(R, d, rng):
    N = d * d
    xs = [np.ones(N, dtype=np.uint8), np.zeros(N, dtype=np.uint8)]
    covs = (0.05, 0.1, 0.15, 0.25, 0.5, 0.7, 0.85, 0.93, 0.97) if d < 16 else (0.1, 0.5, 0.9)
    for cov in covs:
        xs.append((rng.random(N) < cov).astype(np.uint8))
    one = np.ones(N, dtype=np.uint8)
    one[[1, 2]] = 0                                   # two adjacent vacancies at the cell edge
    xs.append(one)
    if d == 12:
        cat = R.LSC_cat(dimen=12)
        for i in (0, 20, 47):
            R.make_random_structure(i, 48, cat, None)
            xs.append(np.array(cat.variable_occs, dtype=np.uint8))
    return xs


def types_fixture(R):
    out = {}
    rng = np.random.default_rng(2024)
    cell_code = {c: i for i, c in enumerate(ref_exec.Lattice.CELLS)}
    for d in (6, 8, 12, 16):
        t0 = time.time()
        N = d * d
        nh3, lsc = R.NH3_cat(dimen=d), R.LSC_cat(dimen=d)
        xs = structures_for(R, d, rng)
        nh3_t, lsc_t, nh3_exp, lsc_exp, lsc_sites, pairs, cells, rates, srate = [], [], [], [], [], [], [], [], []
        for x in xs:
            nh3.assign_occs([int(v) for v in x])
            nh3.graph_to_KMClattice()
            nh3_t.append(R.node_types(nh3))
            nh3_exp.append(list(nh3.KMC_lat.site_type_inds))
            if d <= 12:
                srate.append(lsc.eval_struc_rate_anal([int(v) for v in x]))     # assign + typing + solve
            else:
                lsc.assign_occs([int(v) for v in x])
                lsc.graph_to_KMClattice()
                srate.append(np.sum(lsc.compute_site_rates_lsc()))
            lsc_t.append(R.node_types(lsc))
            lsc_exp.append(list(lsc.KMC_lat.site_type_inds))
            lsc_sites.append(list(lsc.var_ind_kmc_sites))
            pairs.append([v for pr in lsc.KMC_lat.neighbor_list for v in pr])
            cells.append([cell_code[c] for c in lsc.KMC_lat.cell_list])
            rates.append(lsc.compute_site_rates_lsc())
        tag = "d%d_" % d
        out[tag + "x"] = np.array(xs, dtype=np.uint8)
        out[tag + "nh3_types"] = np.array(nh3_t, dtype=np.uint8)
        out[tag + "lsc_types"] = np.array(lsc_t, dtype=np.uint8)
        out[tag + "nh3_export_ptr"], out[tag + "nh3_export"] = ragged(nh3_exp, np.uint8)
        out[tag + "lsc_export_ptr"], out[tag + "lsc_export"] = ragged(lsc_exp, np.uint8)
        _, out[tag + "lsc_kmc_sites"] = ragged(lsc_sites)
        out[tag + "lsc_pairs_ptr"], out[tag + "lsc_pairs"] = ragged(pairs)
        _, out[tag + "lsc_pair_cells"] = ragged(cells, np.uint8)
        out[tag + "lsc_site_rates"] = np.array(rates)
        out[tag + "lsc_struct_rate"] = np.array(srate, dtype=np.float64)
        n9 = int(sum((t == 9).sum() for t in nh3_t))
        print("types d=%d: %d structures, %d type-9 sites, %.1f s" % (d, len(xs), n9, time.time() - t0))
    np.savez_compressed(os.path.join(HERE, "ref_types.npz"), **out)


def flatten_tree(clf):
    t = clf.tree_
    cls = np.asarray(clf.classes_)
    leaf = cls[np.argmax(t.value[:, 0, :], axis=1)]
    return dict(t_feature=t.feature.astype(np.int32), t_thr=t.threshold.astype(np.float64),
                t_left=t.children_left.astype(np.int32), t_right=t.children_right.astype(np.int32),
                t_active=(leaf == 1.0).astype(np.uint8))


def surrogate_sa_fixture(R):
    d, N, n_strucs = 12, 144, 48
    out = {}
    cat = R.LSC_cat(dimen=d)
    t0 = time.time()
    occs, rates, syms = [], [], []
    for i in range(n_strucs):
        R.make_random_structure(i, n_strucs, cat, None)       # leaves occupancies, graph and KMC lattice on `cat`
        occs.append(list(cat.variable_occs))
        rates.append(cat.compute_site_rates_lsc())
        syms.append(cat.generate_all_translations_and_rotations())
    occs, rates, all_syms = np.array(occs, dtype=np.uint8), np.array(rates), np.vstack(syms)
    print("48 structures + labels + symmetries: %.1f s" % (time.time() - t0))
    t0 = time.time()
    sur = R.train_surrogate(all_syms, rates, seed=0)
    print("reference training recipe: %.1f s, %d epochs, tree nodes %d, active rows %d" % (
        time.time() - t0, sur.predictor.n_iter_, sur.classifier.tree_.node_count, len(sur.Y_reg)))
    out.update(train_occs=occs, train_rates=rates, site_is_active=sur.site_is_active.astype(np.uint8),
               Y_reg=sur.Y_reg, X_reg_rows=np.nonzero(sur.site_is_active)[0].astype(np.int32),
               W1=sur.predictor.coefs_[0], b1=sur.predictor.intercepts_[0], w2=sur.predictor.coefs_[1].reshape(-1),
               b2=sur.predictor.intercepts_[1], y_mean=sur.Y_scaler.mean_[0], y_scale=sur.Y_scaler.scale_[0],
               loss_curve=np.array(sur.predictor.loss_curve_), n_iter=np.int64(sur.predictor.n_iter_),
               clf_train_pred=sur.classifier.predict(all_syms[::7]).astype(np.uint8), **flatten_tree(sur.classifier))
    assert (all_syms[sur.site_is_active == 1] == sur.X_reg).all()

    # an always-active classifier ("gate off"): the same sklearn class fitted on all-ones labels
    from sklearn import tree as sktree
    open_clf = sktree.DecisionTreeClassifier(max_depth=30, random_state=0).fit(all_syms[:50], np.ones(50))

    # scores: eval_structure_rate on training structures and random ones
    rng = np.random.default_rng(77)
    score_x = np.vstack([occs[::6], (rng.random((8, N)) < rng.random((8, 1))).astype(np.uint8)])
    sc, scn, sco = [], [], []
    for x in score_x:
        cat.variable_occs = [int(v) for v in x]
        T = cat.generate_all_translations()
        sc.append(sur.eval_rate(T))
        scn.append(sur.eval_structure_rate(T, normalize=True))
        gate = sur.classifier
        sur.classifier = open_clf
        sco.append(sur.eval_rate(T))
        sur.classifier = gate
    out.update(score_x=score_x, score_gated=np.array(sc, dtype=np.float64), score_gated_norm=np.array(scn, dtype=np.float64),
               score_open=np.array(sco, dtype=np.float64))

    # SA runs: injected draws from the generator BASELINE config 2 names
    g = np.random.Generator(np.random.Philox(key=1234))
    n_chains, max_steps = 4, 14400
    x0 = (g.random((n_chains, N)) < 0.5).astype(np.uint8)
    prop = g.integers(0, N, size=(n_chains, max_steps)).astype(np.int32)
    u = g.random((n_chains, max_steps), dtype=np.float32)
    out.update(sa_x0=x0, sa_prop=prop, sa_u=u, T_hot=np.float64(rates.max()))
    runs = []          # (name, chain, gated, n_cycles, T_0, cooling)
    for c in range(n_chains):
        runs += [("gated_cold_c%d" % c, c, True, 10, 0.05, "exp"), ("gated_hot_c%d" % c, c, True, 10, float(rates.max()), "exp"),
                 ("open_cold_c%d" % c, c, False, 10, 0.05, "exp"), ("open_hot_c%d" % c, c, False, 10, float(rates.max()), "exp")]
    runs += [("gated_log", 0, True, 3, 0.05, "log"), ("gated_linear", 1, True, 3, 0.05, "linear"),
             ("gated_quench", 2, True, 3, 0.05, "quench"), ("config1_full", 0, True, 100, 0.05, "exp")]
    names = []
    for name, c, gated, n_cycles, T_0, cooling in runs:
        t0 = time.time()
        steps = n_cycles * N
        gate = sur.classifier
        if not gated:
            sur.classifier = open_clf
        r = R.run_optimize(cat, x0[c], prop[c, :steps], u[c, :steps], surrogate=sur, n_cycles=n_cycles, T_0=T_0,
                           cooling=cooling)
        sur.classifier = gate
        assert len(r["accept"]) == steps
        out["sa_%s_x" % name] = r["x"]
        out["sa_%s_accept" % name] = np.packbits(r["accept"])
        out["sa_%s_OF_new" % name] = r["OF_new"]
        out["sa_%s_OF0" % name] = np.float64(r["OF0"])
        out["sa_%s_traj" % name] = r["traj"]
        out["sa_%s_u_used" % name] = r["u_used"].astype(np.int32)
        out["sa_%s_meta" % name] = np.array([c, int(gated), n_cycles, T_0, ["exp", "log", "linear", "quench"].index(cooling)],
                                            dtype=np.float64)
        names.append(name)
        print("optimize() %-16s %6d steps  accept %.3f  %.1f s" % (name, steps, r["accept"].mean(), time.time() - t0))
    out["sa_names"] = np.array(names)

    # mode='analytical' (LSC): d = 6 and d = 12
    for dd, n_cycles in ((6, 8), (12, 1)):
        lsc = R.LSC_cat(dimen=dd)
        NN = dd * dd
        gg = np.random.default_rng(50 + dd)
        xa = (gg.random(NN) < 0.8).astype(np.uint8)
        pa = gg.integers(0, NN, n_cycles * NN).astype(np.int32)
        ua = gg.random(n_cycles * NN, dtype=np.float32)
        t0 = time.time()
        r = R.run_optimize(lsc, xa, pa, ua, mode="analytical", n_cycles=n_cycles, T_0=0.05, cooling="exp")
        tag = "anal_d%d_" % dd
        out.update({tag + "x0": xa, tag + "prop": pa, tag + "u": ua, tag + "x": r["x"], tag + "accept": r["accept"],
                    tag + "OF_new": r["OF_new"], tag + "OF0": np.float64(r["OF0"]), tag + "traj": r["traj"]})
        print("optimize(analytical) d=%d: %d steps, accept %.3f, %.1f s" % (dd, len(pa), r["accept"].mean(), time.time() - t0))
    np.savez_compressed(os.path.join(HERE, "ref_surrogate_sa.npz"), **out)


if __name__ == "__main__":
    if not ref_exec.available():
        sys.exit("needs the reference sources under " + ref_exec.REF_ROOT)
    R = ref_exec.reference()
    which = sys.argv[1:] or ["dynamic_cat", "types", "surrogate_sa"]
    if "dynamic_cat" in which:
        dynamic_cat_fixture(R)
    if "types" in which:
        types_fixture(R)
    if "surrogate_sa" in which:
        surrogate_sa_fixture(R)
    for f in sorted(os.listdir(HERE)):
        if f.startswith("ref_"):
            print("%-28s %8d bytes" % (f, os.path.getsize(os.path.join(HERE, f))))
