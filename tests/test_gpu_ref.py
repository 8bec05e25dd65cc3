"""GPU: the product (reference-facing Python API and raw engine calls) against fixtures produced by EXECUTING THE
REFERENCE'S OWN CODE (tests/golden/ref_*.npz, see tests/ref_exec.py / tests/golden/make_ref_golden.py).  Bit-exact for
occupancies, site types, accept/reject vectors and index maps; 1e-5 relative for objectives (stated per assert)."""
import os
import random
import numpy as np
import pytest

from oracle import lattice as L
import helpers as Hp
from test_gpu_api import duck_surrogate
from test_ref_golden import ref_surrogate, _sa_runs

pytestmark = pytest.mark.gpu


def _load(name):
    return np.load(os.path.join(Hp.GOLDEN, name))


@pytest.mark.parametrize("d", [6, 12])
def test_dynamic_cat_api_equals_reference_execution(d):
    """OML/dynamic_cat.py:62-84,172-363,188-220,404-437 executed upstream vs the drop-in classes."""
    from so_b200 import LSC_cat, dynamic_cat, dyn_cat_dist
    z = _load("ref_dynamic_cat.npz")
    t = "d%d_" % d
    N = d * d
    cat = LSC_cat(d)
    x = [int(v) for v in z[t + "x"]]
    cat.variable_occs = list(x)
    for (a, b), row in zip(z[t + "shifts"], z[t + "translate"]):
        np.testing.assert_array_equal(cat.translate(int(a), int(b)), row)
    for k, row in zip((0, 1, 2, 4), z[t + "rotate"]):
        np.testing.assert_array_equal(np.asarray(cat.rotate(k), dtype=float), row)
    for (a, b), row in zip(z[t + "shifts"], z[t + "local_inds"]):
        assert cat.get_local_inds(int(a), int(b)) == row.tolist()
    assert [cat.var_ind_to_sym_inds(v) for v in range(N)] == [tuple(r) for r in z[t + "var2sym"].tolist()]
    assert [cat.sym_inds_to_var_ind(int(a), int(b)) for a, b in z[t + "sym_in"]] == z[t + "sym2var"].tolist()
    if d == 6:
        np.testing.assert_array_equal(cat.generate_all_translations(), z[t + "all_trans"])
        np.testing.assert_array_equal(cat.generate_all_translations_and_rotations(), z[t + "all_syms"])
    else:
        np.testing.assert_array_equal(cat.generate_all_translations()[[0, 1, 13, 77, 143]], z[t + "all_trans_rows"])
        np.testing.assert_array_equal(cat.generate_all_translations_and_rotations()[[0, 1, 2, 40, 41, 431]],
                                      z[t + "all_syms_rows"])
    assert list(cat.variable_atoms) == z[t + "variable_atoms"].tolist()
    assert abs(cat.surface_area - float(z[t + "surface_area"])) < 1e-9
    # the global `random` stream is consumed exactly as upstream does
    random.seed(7)
    assert cat.randomize(coverage=0.5) == z[t + "randomize_half"].tolist()
    assert cat.randomize() == z[t + "randomize_none"].tolist()
    assert cat.randomize(coverage=1.0 / 3) == z[t + "randomize_third"].tolist()
    random.seed(11)
    cat.variable_occs = list(x)
    for k in range(12):
        cat.rand_move()
        assert cat.atom_last_moved == int(z[t + "moved"][k])
        if k % 3 == 2:
            cat.revert_last()
        assert cat.variable_occs == z[t + "move_states"][k].tolist()
    cat.flip_atom(3 * N + 5)
    assert cat.variable_occs == z[t + "move_states"][12].tolist()
    # GA helpers (geo_crossover consumes random.random() pt1 + pt2 times)
    random.seed(5)
    x2 = [int(v) for v in z[t + "xo_parent2"]]
    c1, c2 = cat.geo_crossover(list(x), list(x2))
    c3, c4 = cat.geo_crossover(list(x), list(x2), pt1=2, pt2=3)
    assert [c1, c2, c3, c4] == z[t + "xo_children"].tolist()
    cat.variable_occs = list(x)
    w = cat.calc_weights()
    np.testing.assert_allclose(w, z[t + "weights"], rtol=1e-15)
    assert abs(dyn_cat_dist(np.array(x, dtype=float), np.array(x2, dtype=float), w) - float(z[t + "dist"])) < 1e-12
    if d == 6:
        base = dynamic_cat(dim1=6, dim2=6)
        assert list(base.variable_atoms) == z["base_variable_atoms"].tolist()
        assert base.variable_occs == z["base_occs"].tolist()
        assert abs(base.surface_area - float(z["base_surface_area"])) < 1e-9


@pytest.mark.parametrize("d", [6, 8, 12, 16])
def test_site_typing_equals_reference_execution(d):
    """graph_to_KMClattice() of both catalysts executed upstream vs descriptors_kernel / the drop-in classes."""
    from so_b200 import engine, LSC_cat, NH3_cat
    z = _load("ref_types.npz")
    t = "d%d_" % d
    xs = z[t + "x"]
    N = d * d
    n = len(xs)
    # raw engine, all structures in one launch: all 19 NH3 types on 5N nodes, LSC types, Ni-neighbour counts
    lat = engine.Lattice(d, "NH3")
    ch = engine.Chains(lat, n)
    ch.set_occupancy(xs)
    out = ch.descriptors(c6=True, lsc_type=True, nh3_type=True, lsc_hist=True, nh3_hist=True, nh3_exp=True)
    assert (out["nh3_type"] == z[t + "nh3_types"]).all()
    assert (out["lsc_type"] == z[t + "lsc_types"][:, 3 * N:4 * N]).all()
    for i in range(n):
        assert (out["nh3_hist"][i] == np.bincount(z[t + "nh3_types"][i], minlength=20)).all()
        ref_lsc = z[t + "lsc_types"][i, 3 * N:4 * N]
        assert out["lsc_hist"][i, 1] == (ref_lsc == 1).sum() and out["lsc_hist"][i, 2] == (ref_lsc == 2).sum()
    ch.close()
    # drop-in classes: KMC_lat exports, var_ind_kmc_sites, neighbour pairs, analytical rates
    nh3, lsc = NH3_cat(d), LSC_cat(d)
    ep, ev, lp, lv, ls = (z[t + k] for k in ("nh3_export_ptr", "nh3_export", "lsc_export_ptr", "lsc_export", "lsc_kmc_sites"))
    pp, pv, pc = z[t + "lsc_pairs_ptr"], z[t + "lsc_pairs"], z[t + "lsc_pair_cells"]
    cells = ("self", "north", "northeast", "east", "southeast")
    for i in range(0, n, 2 if d >= 12 else 1):
        x = [int(v) for v in xs[i]]
        nh3.assign_occs(list(x))
        nh3.graph_to_KMClattice()
        assert nh3.KMC_lat.site_type_inds == ev[ep[i]:ep[i + 1]].tolist()
        got = np.array([nh3.defected_graph.node[k]['site_type'] or 0 for k in range(5 * N)])
        assert (got == z[t + "nh3_types"][i]).all()
        lsc.assign_occs(list(x))
        lsc.graph_to_KMClattice()
        assert lsc.KMC_lat.site_type_inds == lv[lp[i]:lp[i + 1]].tolist()
        assert lsc.var_ind_kmc_sites == ls[lp[i]:lp[i + 1]].tolist()
        ref_pairs = [tuple(p) + (cells[c],) for p, c in zip(pv[pp[i]:pp[i + 1]].reshape(-1, 2).tolist(), pc[pp[i] // 2:pp[i + 1] // 2])]
        got_pairs = [tuple(p) + (c,) for p, c in zip(lsc.KMC_lat.neighbor_list, lsc.KMC_lat.cell_list)]
        assert sorted(got_pairs) == sorted(ref_pairs)
        if d <= 12:
            rates = lsc.compute_site_rates_lsc() if lsc.KMC_lat.site_type_inds else np.zeros(N)
            np.testing.assert_allclose(rates, z[t + "lsc_site_rates"][i], rtol=1e-9, atol=1e-13)
            np.testing.assert_allclose(lsc.eval_struc_rate_anal(list(x)), z[t + "lsc_struct_rate"][i], rtol=1e-9, atol=1e-13)


def test_scores_equal_reference_execution():
    """eval_structure_rate (OML/train_surrogate.py:34-54) executed upstream with its sklearn models vs the GPU scorers:
    1e-5 relative (the north_star bar for fp32-class arithmetic)."""
    from so_b200 import engine, LSC_cat, surrogate
    z = _load("ref_surrogate_sa.npz")
    d = 12
    lat = engine.Lattice(d, "LSC")
    for gated, key in ((True, "score_gated"), (False, "score_open")):
        p = ref_surrogate(z, gated)
        tab = Hp.tables_from_params(lat, p)
        xs = z["score_x"]
        ch = engine.Chains(lat, len(xs))
        ch.set_occupancy(xs)
        ch.attach(tab)
        of = ch.objective()
        rate, _, _ = tab.score_batch(engine.pack_bits(xs))
        for a, b, ref in zip(of, rate, z[key]):
            assert abs(a - ref) <= 1e-5 * max(abs(ref), 1e-12) and abs(b - ref) <= 1e-5 * max(abs(ref), 1e-12)
        ch.close()
    sur = surrogate()
    dk = duck_surrogate(ref_surrogate(z, True))
    sur.predictor, sur.classifier, sur.Y_scaler = dk.predictor, dk.classifier, dk.Y_scaler
    sur.cat = LSC_cat(d)
    for x, sg, sn in zip(z["score_x"], z["score_gated"], z["score_gated_norm"]):
        T = L.generate_all_translations(x, d)
        assert abs(sur.eval_rate(T) - sg) <= 1e-5 * max(abs(sg), 1e-12)
        assert abs(sur.eval_structure_rate(T, normalize=True) - sn) <= 1e-5 * max(abs(sn), 1e-12)


def test_partition_and_device_training_equal_reference_execution():
    """partition_data_set + train_regressor (OML/train_surrogate.py:72-98,141-152) executed upstream (global NumPy
    stream seeded with 0, as MLPRegressor(random_state=None) consumes it) vs the drop-in with the device trainer."""
    from so_b200 import surrogate
    z = _load("ref_surrogate_sa.npz")
    d = 12
    all_syms = np.vstack([L.generate_all_translations_and_rotations(x, d) for x in z["train_occs"]])
    sur = surrogate()
    sur.all_syms = all_syms
    sur.partition_data_set(z["train_rates"])
    assert (sur.site_is_active == z["site_is_active"]).all()
    np.testing.assert_array_equal(sur.Y_reg, z["Y_reg"])
    assert (sur.X_reg == all_syms[z["X_reg_rows"]]).all()
    sur.train_classifier()
    assert (sur.classifier.predict(all_syms[::7]) == z["clf_train_pred"]).all()
    np.random.seed(0)
    sur.train_regressor(max_iter=10000, random_state=None, on_device=True)
    assert sur.predictor.n_iter_ == int(z["n_iter"])                           # same stopping epoch
    np.testing.assert_allclose(sur.predictor.loss_curve_, z["loss_curve"], rtol=1e-7)
    np.testing.assert_allclose(sur.predictor.coefs_[0], z["W1"], rtol=1e-5, atol=1e-7)
    assert abs(sur.Y_scaler.mean_[0] - float(z["y_mean"])) < 1e-14 and abs(sur.Y_scaler.scale_[0] - float(z["y_scale"])) < 1e-14


def test_optimize_equals_reference_execution():
    """optimize() (OML/optimize_SA.py:11-138) executed upstream with injected draws vs the drop-in optimize(): final
    occupancy left on the catalyst and the recorded trajectory, for every run of the fixture (gated / always-active
    classifier, cold / hot T_0, exp / log / linear(py2) / quench, and the full 14,400-step config-1 run)."""
    from so_b200 import LSC_cat, optimize
    z = _load("ref_surrogate_sa.npz")
    N = 144
    for name, c, gated, n_cycles, T_0, cooling in _sa_runs(z):
        steps = n_cycles * N
        sur = duck_surrogate(ref_surrogate(z, gated))
        cat = LSC_cat(12)
        cat.assign_occs([int(v) for v in z["sa_x0"][c]])
        py2 = cooling == "linear_py2"
        out = optimize(cat, surrogate=sur, n_cycles=n_cycles, T_0=T_0, cooling="linear" if py2 else cooling,
                       py2_linear=py2, n_record=100, proposals=z["sa_prop"][c, :steps], uniforms=z["sa_u"][c, :steps],
                       return_stats=True)
        assert cat.variable_occs == z["sa_%s_x" % name].tolist(), name
        acc = np.unpackbits(z["sa_%s_accept" % name])[:steps]
        assert out["accepted"] == int(acc.sum()) and out["flips"] == steps
        traj = z["sa_%s_traj" % name]
        np.testing.assert_array_equal(out["trajectory"][0], traj[0])
        scale = max(np.abs(traj[1]).max(), 1e-12)
        assert np.abs(out["trajectory"][1] - traj[1]).max() <= 1e-5 * scale


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5])
def test_accept_vectors_equal_reference_execution(variant):
    """so_anneal's accept/reject bytes vs the reference-executed optimize(), every kernel variant that can run the
    full 144-column descriptor: the chains of one (gate, T_0) group share a launch."""
    from so_b200 import engine
    z = _load("ref_surrogate_sa.npz")
    d, N = 12, 144
    lat = engine.Lattice(d, "LSC")
    groups = {}
    for run in _sa_runs(z):
        name, c, gated, n_cycles, T_0, cooling = run
        if n_cycles == 10:
            groups.setdefault((gated, T_0), []).append(run)
    assert len(groups) == 4
    for (gated, T_0), runs in groups.items():
        if variant == 1 and not gated:
            continue                                    # variant 1 is the row-per-lane gated kernel
        p = ref_surrogate(z, gated)
        tab = Hp.tables_from_params(lat, p)
        steps = 10 * N
        cs = [r[1] for r in runs]
        ch = engine.Chains(lat, len(cs))
        ch.set_occupancy(z["sa_x0"][cs])
        ch.attach(tab)
        temps = T_0 * np.exp(-np.arange(steps, dtype=np.float64) / (steps / 5.0))
        w2sum = float(np.abs(p.w2).sum())
        tol = 8.0 * np.sqrt(tab.K) * np.ldexp(1.0, tab.e_h) * abs(p.y_scale) * w2sum
        st, acc = ch.anneal(temps, proposals=z["sa_prop"][cs, :steps], uniforms=z["sa_u"][cs, :steps], want_accept=True,
                            exact_guard=True, guard_tol=tol, variant=variant)
        x1 = ch.get_occupancy()
        for k, (name, c, _, _, _, _) in enumerate(runs):
            ref_acc = np.unpackbits(z["sa_%s_accept" % name])[:steps]
            assert (acc[k] == ref_acc).all(), (name, variant)
            assert (x1[k] == z["sa_%s_x" % name]).all()
        ch.close()


@pytest.mark.parametrize("d", [6, 12])
def test_analytical_mode_equals_reference_execution(d):
    """optimize(mode='analytical') (OML/optimize_SA.py:43-44,103-104 + OML/LSC_cat.py:248-316) executed upstream vs
    so_anneal_analytical."""
    from so_b200 import LSC_cat, optimize
    z = _load("ref_surrogate_sa.npz")
    t = "anal_d%d_" % d
    cat = LSC_cat(d)
    cat.assign_occs([int(v) for v in z[t + "x0"]])
    n_cycles = len(z[t + "prop"]) // (d * d)
    out = optimize(cat, mode='analytical', n_cycles=n_cycles, T_0=0.05, cooling='exp', proposals=z[t + "prop"],
                   uniforms=z[t + "u"], return_stats=True)
    assert cat.variable_occs == z[t + "x"].tolist()
    assert out["accepted"] == int(z[t + "accept"].sum())
    traj = z[t + "traj"]
    np.testing.assert_allclose(out["trajectory"][1], traj[1], rtol=1e-9, atol=1e-13)


def test_island_generator_equals_reference_execution():
    """drivers/generate_init_structures.py make_random_structure executed upstream (all 48 structures of the fixture)
    vs the device generator (so_island_structures: NumPy's legacy MT19937 stream restated in the kernel) and the host
    recipe of so_b200.database."""
    from so_b200 import engine, database, LSC_cat
    z = _load("ref_surrogate_sa.npz")
    lat = engine.Lattice(12, "LSC")
    got = database.make_random_structures_device(lat, np.arange(48), 48)
    assert (got == z["train_occs"]).all()
    cat = LSC_cat(12)
    for i in (0, 13, 47):
        assert database.make_random_structure(i, 48, cat) == z["train_occs"][i].tolist()
    # other sizes / counts: against NumPy's own stream
    for d, n_strucs, idx in ((20, 10, [0, 3, 9]), (96, 7, [2, 6])):
        N = d * d
        lat = engine.Lattice(d, "LSC")
        got = database.make_random_structures_device(lat, idx, n_strucs)
        for row, i in zip(got, idx):
            np.random.seed(i)
            x = np.zeros(N, dtype=np.uint8)
            for t in np.random.choice(range(N), size=int(3 + 15 * float(i) / (n_strucs + 1)), replace=False):
                for a, b in database.ISLAND:
                    x[((t // d + b) % d) * d + (t % d + a) % d] = 1
            for s in np.random.choice(range(N), size=16, replace=False):
                x[s] ^= 1
            assert (row == x).all()


def test_device_classifier_and_training_set_equal_reference_execution():
    """Rest of f-2 on the device: all_syms (x3 rotation augmentation) and the activity labels from so_build_training_set,
    the depth-30 CART from so_tree_fit -- against generate_all_translations_and_rotations / partition_data_set /
    train_classifier EXECUTED FROM THE REFERENCE SOURCE (the tree of tests/golden/ref_surrogate_sa.npz, node for node)."""
    from so_b200 import engine, surrogate, LSC_cat
    from so_b200.train_surrogate import build_training_set, DeviceTreeClassifier, flatten_tree
    from oracle import tree as OT
    z = _load("ref_surrogate_sa.npz")
    d = 12
    cat = LSC_cat(d)
    X, active, y_max = build_training_set(cat._lattice, z["train_occs"], z["train_rates"])
    want = np.vstack([L.generate_all_translations_and_rotations(x, d) for x in z["train_occs"][:3]])
    assert (X[:len(want)] == want).all() and X.shape == (48 * 3 * 144, 144)
    assert (active == z["site_is_active"]).all() and y_max == z["train_rates"].max()
    zd = _load("ref_dynamic_cat.npz")
    X6, _, _ = build_training_set(engine.Lattice(6, "LSC"), zd["d6_x"][None, :])
    assert (X6 == zd["d6_all_syms"]).all()                         # the reference's own 108 x 36 matrix
    sur = surrogate()
    sur.cat = cat
    sur.all_syms = X
    sur.partition_data_set(z["train_rates"])
    assert (sur.site_is_active == z["site_is_active"]).all() and (sur.Y_reg == z["Y_reg"]).all()
    sur.train_classifier()                                          # DeviceTreeClassifier
    assert isinstance(sur.classifier, DeviceTreeClassifier)
    flat = flatten_tree(sur.classifier)
    for k, zk in (("feature", "t_feature"), ("thr", "t_thr"), ("left", "t_left"), ("right", "t_right"), ("active", "t_active")):
        assert (flat[k] == z[zk]).all(), k
    assert (sur.classifier.predict(X[::7]) == z["clf_train_pred"]).all()
    # random 0/1 problems: the device tree equals the oracle's restatement of scikit-learn (which tests/test_oracle_tree.py
    # pins to the installed scikit-learn) -- constant and duplicated columns, shallow depth, one class
    rng = np.random.default_rng(5)
    for case in range(6):
        n, K = int(rng.integers(30, 3000)), int(rng.integers(3, 200))
        Xr = (rng.random((n, K)) < rng.random()).astype(np.float64)
        Xr[:, 0] = 1.0
        Xr[:, K - 1] = Xr[:, 1]
        y = ((Xr @ rng.normal(size=K) + 0.3 * rng.normal(size=n)) > 0).astype(np.float64) if case < 5 else np.ones(n)
        depth = (30, 3, 30, 8, 30, 30)[case]
        t = OT.fit_tree(Xr, y, max_depth=depth, random_state=case % 2)
        clf = DeviceTreeClassifier(max_depth=depth, random_state=case % 2).fit(Xr, y)
        assert (clf.tree_.feature == t["feature"]).all() and (clf.tree_.children_left == t["children_left"]).all()
        assert (clf.tree_.children_right == t["children_right"]).all() and (clf.tree_.threshold == t["threshold"]).all()
        assert (clf.tree_.n_node_samples == t["count0"] + t["count1"]).all()
