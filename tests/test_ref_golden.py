"""CPU: the oracle against fixtures produced by EXECUTING THE REFERENCE'S OWN CODE (tests/golden/ref_*.npz, written by
tests/golden/make_ref_golden.py through tests/ref_exec.py).  This is what pins the oracle: site typing
(OML/LSC_cat.py:164-245, OML/NH3_cat.py:158-428), index maps / translations / rotations (OML/dynamic_cat.py:233-363),
the surrogate forward (OML/train_surrogate.py:34-54), the training recipe (:72-152), the SA loop
(OML/optimize_SA.py:11-138) and the analytical objective (OML/LSC_cat.py:248-316).

The *_live tests re-execute the reference in this container (skipped where /root/reference is absent) to show the
committed fixtures are what the reference produces today."""
import os
import random
import numpy as np
import pytest
from oracle import lattice as L, site_types as ST, surrogate as S, sa as SA, lsc_analytical as LA
import helpers as Hp
import ref_exec


def _load(name):
    return np.load(os.path.join(Hp.GOLDEN, name))


# ---- OML/dynamic_cat.py -----------------------------------------------------------------------------------
@pytest.mark.parametrize("d", [6, 12])
def test_index_maps_translations_rotations(d):
    z = _load("ref_dynamic_cat.npz")
    t = "d%d_" % d
    x = z[t + "x"]
    N = d * d
    for (a, b), row in zip(z[t + "shifts"], z[t + "translate"]):
        assert (L.translate(x, int(a), int(b), d) == row).all()
    for k, row in zip((0, 1, 2, 4), z[t + "rotate"]):
        assert (L.rotate(x, k, d) == row).all()
    for (a, b), row in zip(z[t + "shifts"], z[t + "local_inds"]):
        assert L.get_local_inds(int(a), int(b), d) == row.tolist()
    assert [L.var_ind_to_sym_inds(v, d) for v in range(N)] == [tuple(r) for r in z[t + "var2sym"].tolist()]
    assert [L.sym_inds_to_var_ind(int(a), int(b), d) for a, b in z[t + "sym_in"]] == z[t + "sym2var"].tolist()
    if d == 6:
        assert (L.generate_all_translations(x, d) == z[t + "all_trans"]).all()
        assert (L.generate_all_translations_and_rotations(x, d) == z[t + "all_syms"]).all()
    else:
        assert (L.generate_all_translations(x, d)[[0, 1, 13, 77, 143]] == z[t + "all_trans_rows"]).all()
        assert (L.generate_all_translations_and_rotations(x, d)[[0, 1, 2, 40, 41, 431]] == z[t + "all_syms_rows"]).all()
    assert z[t + "variable_atoms"].tolist() == list(range(3 * N, 4 * N))
    assert abs(float(z[t + "surface_area"]) - abs(np.linalg.det(L.cell_vectors(d)))) < 1e-9
    # local descriptor column order == get_local_inds order
    off = L.local_offsets(3)
    assert L.get_local_inds(2, 5, d) == [L.sym_inds_to_var_ind(2 + int(a), 5 + int(b), d) for a, b in off]


# ---- site typing ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d", [6, 8, 12, 16])
def test_site_types_equal_reference_execution(d):
    z = _load("ref_types.npz")
    t = "d%d_" % d
    xs = z[t + "x"]
    N = d * d
    nh3 = ST.nh3_types(xs, d)
    assert (nh3 == z[t + "nh3_types"]).all()                       # all 19 types on all 5N nodes, incl. type 9
    lsc = ST.lsc_types(xs, d)
    ref_lsc = z[t + "lsc_types"]
    assert (lsc == ref_lsc[:, 3 * N:4 * N]).all() and ref_lsc[:, :3 * N].sum() == 0 and ref_lsc[:, 4 * N:].sum() == 0
    assert (ST.c6_counts(xs, d) <= 6).all()
    ep, ev, lp, lv, ls = (z[t + k] for k in ("nh3_export_ptr", "nh3_export", "lsc_export_ptr", "lsc_export", "lsc_kmc_sites"))
    pp, pv, pc = z[t + "lsc_pairs_ptr"], z[t + "lsc_pairs"], z[t + "lsc_pair_cells"]
    for i, x in enumerate(xs):
        assert ST.nh3_export(x, d) == ev[ep[i]:ep[i + 1]].tolist()
        types_, sites = ST.lsc_export(x, d)
        assert types_ == lv[lp[i]:lp[i + 1]].tolist() and sites == ls[lp[i]:lp[i + 1]].tolist()
        ref_pairs = pv[pp[i]:pp[i + 1]].reshape(-1, 2)
        assert sorted(tuple(sorted(p)) for p in ref_pairs.tolist()) == sorted(LA.kmc_neighbor_pairs(sites, d))
        np.testing.assert_allclose(LA.compute_site_rates_lsc(x, d), z[t + "lsc_site_rates"][i], rtol=1e-11, atol=1e-14)
        np.testing.assert_allclose(LA.eval_struc_rate_anal(x, d), z[t + "lsc_struct_rate"][i], rtol=1e-11, atol=1e-14)
    assert int((z[t + "nh3_types"] == 9).sum()) > 0                 # the non-periodic rule is exercised at every size


# ---- surrogate ---------------------------------------------------------------------------------------------
def ref_surrogate(z, gated=True):
    tree = dict(feature=z["t_feature"], thr=z["t_thr"], left=z["t_left"], right=z["t_right"], active=z["t_active"])
    return S.SurrogateParams(L.full_offsets(12), z["W1"], z["b1"], z["w2"], z["b2"], float(z["y_mean"]),
                             float(z["y_scale"]), tree if gated else None)


def test_generator_labels_partition_and_scores():
    z = _load("ref_surrogate_sa.npz")
    d = 12
    occs, rates = z["train_occs"], z["train_rates"]
    for i in range(0, 48, 5):
        assert (LA.make_random_structure(i, 48, d) == occs[i]).all()
        np.testing.assert_allclose(LA.compute_site_rates_lsc(occs[i], d), rates[i], rtol=1e-11, atol=1e-14)
    # partition_data_set (train_surrogate.py:72-98): x3 tiling, 0.06*max threshold
    flat = np.transpose(np.tile(rates.flatten(), [3, 1])).flatten()
    active = flat > 0.06 * flat.max()
    assert (active == z["site_is_active"].astype(bool)).all() and (np.nonzero(active)[0] == z["X_reg_rows"]).all()
    np.testing.assert_array_equal(flat[active], z["Y_reg"])
    assert abs(float(z["y_mean"]) - z["Y_reg"].mean()) < 1e-12 and abs(float(z["y_scale"]) - z["Y_reg"].std()) < 1e-12
    # eval_structure_rate (train_surrogate.py:34-54), gated / normalised / always-active classifier
    p, po = ref_surrogate(z, True), ref_surrogate(z, False)
    q = S.quantize(p)
    for x, sg, sn, so in zip(z["score_x"], z["score_gated"], z["score_gated_norm"], z["score_open"]):
        T = L.generate_all_translations(x, d)
        assert abs(S.eval_rate_fp64(p, T) - sg) <= 1e-11 * max(1.0, abs(sg))
        assert abs(S.eval_rate_fp64(p, T, normalize=True) - sn) <= 1e-11 * max(1.0, abs(sn))
        assert abs(S.eval_rate_fp64(po, T) - so) <= 1e-11 * max(1.0, abs(so))
        of = S.score_fixed(p, q, x, d)[0]
        assert abs(of - sg) <= 1e-5 * max(abs(sg), 1e-12)                   # the CUDA path's arithmetic, fp32 bar
    # the classifier's predictions on every 7th training row
    syms = np.vstack([L.generate_all_translations_and_rotations(x, d) for x in occs])[::7]
    assert (S.tree_predict(p.tree, syms) == z["clf_train_pred"].astype(bool)).all()


def _sa_runs(z):
    for name in z["sa_names"]:
        c, gated, n_cycles, T_0, ci = z["sa_%s_meta" % name]
        yield str(name), int(c), bool(gated), int(n_cycles), float(T_0), ["exp", "log", "linear_py2", "quench"][int(ci)]


def test_sa_loop_equals_reference_execution():
    """optimize() executed from the reference source vs the oracle: literal restatement on every run, fixed-point
    delta form (the CUDA path's arithmetic) on every run; accept vectors, objectives, final occupancies."""
    z = _load("ref_surrogate_sa.npz")
    d, N = 12, 144
    min_margin = np.inf
    for name, c, gated, n_cycles, T_0, cooling in _sa_runs(z):
        steps = n_cycles * N
        p = ref_surrogate(z, gated)
        q = S.quantize(p)
        acc = np.unpackbits(z["sa_%s_accept" % name])[:steps].astype(bool)
        x0, prop, u = z["sa_x0"][c], z["sa_prop"][c, :steps], z["sa_u"][c, :steps]
        py2 = cooling == "linear_py2"
        temps = SA.schedule("linear" if py2 else cooling, T_0, steps, py2_linear=py2)
        b = SA.anneal_fixed(p, q, d, x0, temps, prop, u, margins=True)
        assert (b["accept"] == acc).all(), name
        assert (b["x"] == z["sa_%s_x" % name]).all()
        min_margin = min(min_margin, b["margin"].min())
        # objective after the run and per-step candidates within the fp32 bar of the reference's fp64 numbers
        of_new = z["sa_%s_OF_new" % name]
        of_path = np.concatenate([[float(z["sa_%s_OF0" % name])], of_new])
        cur, ref_delta = of_path[0], np.zeros(steps)
        for s in range(steps):
            ref_delta[s] = of_new[s] - cur
            if acc[s]:
                cur = of_new[s]
        scale = np.abs(of_path).max()
        assert np.abs(b["delta"] - ref_delta).max() <= 1e-5 * scale
        assert abs(b["OF"] - cur) <= 1e-5 * max(abs(cur), 1e-12)
        # u is drawn only on non-improving moves at T > 0 (optimize_SA.py:110-114)
        need_u = np.nonzero((ref_delta <= 0) & (temps > 0))[0]
        assert (z["sa_%s_u_used" % name] == need_u).all()
        if steps <= 1440 and c < 2:
            class Sur(object):
                def eval_rate(self, M, p=p):
                    return S.eval_rate_fp64(p, M)
            a = SA.optimize_faithful(x0, d, Sur(), p.offsets, temps, prop, u)
            assert (a["accept"] == acc).all() and (a["x"] == b["x"]).all()
            np.testing.assert_allclose(a["delta"], ref_delta, rtol=0, atol=1e-10 * scale)
            traj = z["sa_%s_traj" % name]                          # (2, n_record+1): steps, OF/N
            np.testing.assert_allclose(a["traj"][:, 0], traj[0])
            np.testing.assert_allclose(a["traj"][:, 1], traj[1], rtol=0, atol=1e-10 * scale)
    print("minimum Metropolis margin |exp(delta/T) - u| over all reference-executed runs: %.3e" % min_margin)
    assert min_margin > 1e-7


@pytest.mark.parametrize("d", [6, 12])
def test_analytical_mode_equals_reference_execution(d):
    z = _load("ref_surrogate_sa.npz")
    t = "anal_d%d_" % d
    steps = len(z[t + "prop"])
    temps = SA.schedule("exp", 0.05, steps)
    o = SA.optimize_analytical(z[t + "x0"], d, temps, z[t + "prop"], z[t + "u"])
    assert (o["accept"] == z[t + "accept"]).all() and (o["x"] == z[t + "x"]).all()
    cur = float(z[t + "OF0"])
    for s in range(steps):
        if z[t + "accept"][s]:
            cur = z[t + "OF_new"][s]
    assert abs(o["OF"] - cur) <= 1e-10 * max(1.0, abs(cur))


# ---- live re-execution (build container only) ---------------------------------------------------------------
needs_ref = pytest.mark.skipif(not ref_exec.available(), reason="reference sources not present (GPU box)")


@needs_ref
def test_live_reference_typing_matches_fixture_and_oracle():
    R = ref_exec.reference()
    z = _load("ref_types.npz")
    d = 8
    nh3, lsc = R.NH3_cat(dimen=d), R.LSC_cat(dimen=d)
    for i in (2, 4, 7):
        x = z["d8_x"][i]
        nh3.assign_occs([int(v) for v in x])
        nh3.graph_to_KMClattice()
        assert (R.node_types(nh3) == z["d8_nh3_types"][i]).all()
        assert abs(lsc.eval_struc_rate_anal([int(v) for v in x]) - z["d8_lsc_struct_rate"][i]) < 1e-13
    x = (np.random.default_rng(99).random(d * d) < 0.3).astype(np.uint8)       # a structure not in the fixture
    nh3.assign_occs([int(v) for v in x])
    nh3.graph_to_KMClattice()
    assert (R.node_types(nh3) == ST.nh3_types(x, d)[0]).all()
    assert np.allclose(nh3.atoms_template.get_positions(), L.slab_positions(d))


@needs_ref
def test_live_reference_optimize_matches_fixture():
    R = ref_exec.reference()
    z = _load("ref_surrogate_sa.npz")
    d, N = 12, 144
    from sklearn.neural_network import MLPRegressor
    from sklearn.preprocessing import StandardScaler
    # an sklearn regressor carrying the fixture's weights (one partial_fit call to allocate, then overwrite)
    mlp = MLPRegressor(hidden_layer_sizes=(20,), activation="relu")
    mlp.partial_fit(np.zeros((2, N)), np.zeros(2))
    mlp.coefs_ = [z["W1"].copy(), z["w2"].reshape(-1, 1).copy()]
    mlp.intercepts_ = [z["b1"].copy(), z["b2"].reshape(1).copy()]
    scaler = StandardScaler().fit(z["Y_reg"].reshape(-1, 1))
    p = ref_surrogate(z, True)

    class TreeFromArrays(object):
        def predict(self, X):
            return S.tree_predict(p.tree, np.asarray(X)).astype(float)
    sur = R.surrogate()
    sur.predictor, sur.classifier = mlp, TreeFromArrays()
    R.set_y_scaler(scaler)
    cat = R.LSC_cat(dimen=d)
    steps = 2 * N
    r = R.run_optimize(cat, z["sa_x0"][1], z["sa_prop"][1, :steps], z["sa_u"][1, :steps], surrogate=sur, n_cycles=2, T_0=0.05)
    # same draws, but tau differs from the 10-cycle fixture run -> compare with the oracle on this schedule
    temps = SA.schedule("exp", 0.05, steps)
    b = SA.anneal_fixed(p, S.quantize(p), d, z["sa_x0"][1], temps, z["sa_prop"][1, :steps], z["sa_u"][1, :steps])
    assert (b["accept"] == r["accept"]).all() and (b["x"] == r["x"]).all()
    assert cat.variable_occs == [int(v) for v in r["x"]]                  # optimize() leaves the optimum on the catalyst
