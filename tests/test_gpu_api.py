"""GPU tests of the reference-facing Python API (catalyst classes, surrogate, optimize) -- they read like the tests the
reference would have had: same calls, results compared with the oracle's restatement of the reference."""
import os
import random
import types
import numpy as np
import pytest

from oracle import lattice as L, site_types as ST, surrogate as S, sa as SA, lsc_analytical as LA
import helpers as Hp

pytestmark = pytest.mark.gpu


class _Predictor(object):
    out_activation_ = 'identity'


def duck_surrogate(p):
    """A reference-style fitted surrogate object from plain arrays (what sklearn objects expose)."""
    s = types.SimpleNamespace()
    s.predictor = _Predictor()
    s.predictor.coefs_ = [p.W1, p.w2.reshape(-1, 1)]
    s.predictor.intercepts_ = [p.b1, np.array([p.b2])]
    s.Y_scaler = types.SimpleNamespace(mean_=np.array([p.y_mean]), scale_=np.array([p.y_scale]))
    s.classifier = None
    if p.tree is not None:
        t = p.tree
        n = len(t["feature"])
        value = np.zeros((n, 1, 2))
        value[np.arange(n), 0, t["active"].astype(int)] = 1.0
        s.classifier = types.SimpleNamespace(
            classes_=np.array([0.0, 1.0]),
            tree_=types.SimpleNamespace(feature=t["feature"], threshold=t["thr"], children_left=t["left"],
                                        children_right=t["right"], value=value))
    return s


def test_drop_in_imports():
    import OML                                                    # noqa: F401
    from OML.LSC_cat import LSC_cat
    from OML.NH3_cat import NH3_cat
    from OML.optimize_SA import optimize
    from OML.train_surrogate import surrogate
    from OML.toy_cat import LSC_cat as toy                         # the names the upstream drivers import
    from OML.NiPt_NH3 import NH3_cat as nipt
    assert toy is LSC_cat and nipt is NH3_cat and callable(optimize) and surrogate


def test_dynamic_cat_api():
    from so_b200 import LSC_cat, dynamic_cat
    with pytest.raises(ValueError):
        dynamic_cat(facet='110')
    cat = LSC_cat()
    assert (cat.dim1, cat.dim2, cat.atoms_per_layer) == (12, 12, 144)
    assert list(cat.variable_atoms) == list(range(432, 576)) and cat.variable_occs == [1] * 144
    np.testing.assert_allclose(cat.surface_area, np.linalg.norm(np.cross([33.295, 0, 0], [16.648, 28.835, 0])), rtol=1e-4)
    random.seed(0)
    x = cat.randomize(coverage=0.5)
    assert sum(x) == 72 and x is cat.variable_occs
    d = 12
    assert cat.sym_inds_to_var_ind(-1, 13) == L.sym_inds_to_var_ind(-1, 13, d)
    assert cat.var_ind_to_sym_inds(100) == L.var_ind_to_sym_inds(100, d)
    np.testing.assert_array_equal(cat.translate(3, -2), L.translate(x, 3, -2, d))
    np.testing.assert_array_equal(cat.generate_all_translations(), L.generate_all_translations(x, d))
    np.testing.assert_array_equal(cat.generate_all_translations_and_rotations(),
                                  L.generate_all_translations_and_rotations(x, d))
    for k in (1, 2):
        np.testing.assert_array_equal(cat.rotate(k), L.rotate(x, k, d))
    assert cat.get_local_inds(5, 7) == L.get_local_inds(5, 7, d)
    # flips use ATOM indices (dynamic_cat.py:172-185)
    before = list(cat.variable_occs)
    cat.flip_atom(432 + 17)
    assert cat.variable_occs[17] == 1 - before[17]
    cat.rand_move()
    cat.revert_last()
    assert [a == b for a, b in zip(cat.variable_occs, before)].count(False) == 1
    cat.variable_occs[3] = 7
    with pytest.raises(NameError):
        cat.flip_atom(432 + 3)


def test_lsc_graph_to_kmc_lattice():
    from so_b200 import LSC_cat
    d = 12
    cat = LSC_cat(d)
    for i in (3, 20, 40):
        x = LA.make_random_structure(i, 48, d)
        cat.assign_occs([int(v) for v in x])
        cat.graph_to_KMClattice()
        types_, sites = ST.lsc_export(x, d)
        assert cat.KMC_lat.site_type_inds == types_ and cat.var_ind_kmc_sites == sites
        assert sorted(tuple(sorted(p)) for p in cat.KMC_lat.neighbor_list) == sorted(LA.kmc_neighbor_pairs(sites, d))
        full = ST.lsc_types(x, d)[0]
        for v in range(0, 144, 7):
            assert cat.defected_graph.node[432 + v]['site_type'] == (int(full[v]) or None)
            assert cat.defected_graph.node[432 + v]['element'] == ('Ni' if x[v] else 'vacancy')
        assert cat.defected_graph.node[0]['element'] == 'Pt' and cat.defected_graph.node[700]['element'] == 'vacuum'
        np.testing.assert_allclose(cat.eval_struc_rate_anal([int(v) for v in x]), LA.eval_struc_rate_anal(x, d), rtol=1e-12)
        np.testing.assert_allclose(cat.compute_site_rates_lsc(), LA.compute_site_rates_lsc(x, d), rtol=1e-10, atol=1e-14)
    assert len(list(cat.defected_graph.neighbors(500))) == 12 and cat.defected_graph.number_of_nodes() == 720


def test_lattice_input_writer(tmp_path):
    from so_b200 import LSC_cat
    cat = LSC_cat(12)
    cat.assign_occs([int(v) for v in LA.make_random_structure(5, 48, 12)])
    cat.graph_to_KMClattice()
    cat.KMC_lat.Write_lattice_input(str(tmp_path))
    text = open(os.path.join(str(tmp_path), 'lattice_input.dat')).read()
    assert '33.295 \t0.000' in text and '16.648 \t28.835' in text              # the reference fixture's cell
    assert 'n_cell_sites \t %d' % len(cat.KMC_lat.site_type_inds) in text


def test_nh3_graph_to_kmc_lattice():
    from so_b200 import NH3_cat
    cat = NH3_cat()
    assert cat.dim1 == 16 and cat.atoms_per_layer == 256
    d = 16
    rng = np.random.default_rng(3)
    for cov in (0.3, 0.7, 0.95):
        x = (rng.random(256) < cov).astype(int)
        cat.assign_occs([int(v) for v in x])
        cat.graph_to_KMClattice()
        assert cat.KMC_lat.site_type_inds == ST.nh3_export(x, d)
        full = ST.nh3_types(x, d)[0]
        got = np.array([cat.defected_graph.node[i]['site_type'] or 0 for i in range(5 * 256)])
        assert (got == full).all()
        hist, exp = cat.site_type_histogram()
        assert (hist == ST.nh3_hist(full[None])[0]).all() and (exp == ST.nh3_export_hist(full[None])[0]).all()
        y = x.copy()
        y[37] ^= 1
        assert (cat.flip_delta(3 * 256 + 37) == ST.nh3_export_hist(ST.nh3_types(y, d))[0] - exp).all()


def test_surrogate_eval_rate_matches_sklearn():
    """eval_rate on the GPU vs. the reference evaluate path on scikit-learn objects (train_surrogate.py:34-54)."""
    from so_b200 import LSC_cat, surrogate
    d = 12
    occs = np.array([LA.make_random_structure(i, 48, d) for i in range(0, 48, 6)])
    rates = np.array([LA.compute_site_rates_lsc(x, d) for x in occs])
    all_syms = np.vstack([L.generate_all_translations_and_rotations(x, d) for x in occs])
    sur = surrogate()
    sur.train(all_syms, rates, max_iter=60, random_state=0)
    ref = S.FaithfulSurrogate(sur.predictor, sur.Y_scaler, sur.classifier)
    sur.cat = LSC_cat(d)
    for x in occs[:4]:
        T = L.generate_all_translations(x, d)
        a, b = sur.eval_rate(T), ref.eval_rate(T)
        assert abs(a - b) <= 1e-5 * max(abs(b), 1e-9), (a, b)
        assert abs(sur.eval_structure_rate(T, normalize=True) - b / 144) <= 1e-5 * max(abs(b) / 144, 1e-9)


def test_optimize_golden_chain():
    """optimize() with the injected proposal/uniform sequence of the committed golden run: final occupancy and
    trajectory must be those of the reference algorithm (faithful fp64 oracle)."""
    from so_b200 import LSC_cat, optimize
    z = np.load(os.path.join(Hp.GOLDEN, "sa_d12_golden.npz"))
    for name, gated in (("gated", True), ("open", False)):
        p = Hp.golden_surrogate(gated=gated)
        for c in (0, 3):
            cat = LSC_cat(12)
            cat.assign_occs([int(v) for v in z["x0"][c]])
            out = optimize(cat, surrogate=duck_surrogate(p), n_cycles=10, T_0=0.05, cooling='exp', n_record=100,
                           proposals=z["prop"][c], uniforms=z["u"][c], return_stats=True)
            assert cat.variable_occs == [int(v) for v in z[name + "_faithful_x"][c]]
            traj = z[name + "_faithful_traj"][c]                     # rows (step, OF/N, T)
            np.testing.assert_array_equal(out["trajectory"][0], traj[:, 0])
            np.testing.assert_allclose(out["trajectory"][1], traj[:, 1], rtol=1e-5, atol=1e-9)
            np.testing.assert_allclose(out["temperatures"][1:], traj[1:, 2], rtol=1e-12)
            assert out["flips"] == 1440


def test_optimize_many_chains_and_files(tmp_path):
    from so_b200 import LSC_cat, optimize
    p = Hp.golden_surrogate(gated=True)
    cat = LSC_cat(12)
    random.seed(1)
    cat.randomize(coverage=0.5)
    x0 = list(cat.variable_occs)
    sur = duck_surrogate(p)
    start = S.eval_rate_fp64(p, L.generate_all_translations(np.array(x0), 12))
    out = optimize(cat, surrogate=sur, syms=cat.generate_all_translations(), n_cycles=20, T_0=0.05, n_chains=256,
                   seed=7, fldr=str(tmp_path), prefix='t_', return_stats=True, exact_guard=False)
    final = S.eval_rate_fp64(p, L.generate_all_translations(np.array(cat.variable_occs), 12))
    assert abs(final - out["OF"]) <= 1e-5 * abs(final) and final > start
    traj = np.load(os.path.join(str(tmp_path), 't_sim_anneal_trajectory.npy'))
    assert traj.shape == (2, 101) and traj[0, -1] == 20 * 144
    with pytest.raises(NameError):
        optimize(cat, mode='nonsense', surrogate=sur)
    with pytest.raises(NameError):
        optimize(cat, surrogate=sur, cooling='nonsense', n_cycles=1)
    # same seed, same result (counter-based RNG)
    cat2 = LSC_cat(12)
    cat2.assign_occs(list(x0))
    optimize(cat2, surrogate=sur, n_cycles=20, T_0=0.05, n_chains=256, seed=7, exact_guard=False)
    assert cat2.variable_occs == cat.variable_occs


def test_optimize_parallel_tempering():
    """optimize(..., ladder=R): temperature ladders + swap rounds + best-structure record through the drop-in call
    (single process = world 1; the multi-rank path is scripts/optimize_multi_gpu.py under torchrun)."""
    from so_b200 import LSC_cat, optimize
    p = Hp.golden_surrogate(gated=False)
    sur = duck_surrogate(p)
    random.seed(2)
    outs = []
    for rep in range(2):
        cat = LSC_cat(12)
        random.seed(2)
        cat.randomize(coverage=0.5)
        x0 = np.array(cat.variable_occs)
        out = optimize(cat, surrogate=sur, n_cycles=12, T_0=0.05, n_chains=64, seed=11, ladder=8, gather_every=300,
                       exact_guard=False, return_stats=True)
        outs.append((list(cat.variable_occs), out))
    (xa, a), (xb, b) = outs
    assert xa == xb and a["OF"] == b["OF"] and a["temperature_swaps"] == b["temperature_swaps"] > 0     # deterministic
    final = S.eval_rate_fp64(p, L.generate_all_translations(np.array(xa), 12))
    start = S.eval_rate_fp64(p, L.generate_all_translations(x0, 12))
    assert abs(final - a["OF"]) <= 1e-5 * abs(final) and final > start
    assert a["flips"] == 12 * 144 * 64 and 0 <= a["best_chain"] < 64
    with pytest.raises(ValueError):
        optimize(LSC_cat(12), surrogate=sur, n_cycles=1, n_chains=10, ladder=8)          # chains not a multiple of the ladder


def test_optimize_analytical_mode():
    from so_b200 import LSC_cat, optimize
    cat = LSC_cat(12)
    x = LA.make_random_structure(9, 48, 12)
    cat.assign_occs([int(v) for v in x])
    start = LA.eval_struc_rate_anal(x, 12)
    random.seed(3)
    optimize(cat, mode='analytical', n_cycles=1, T_0=0.0, cooling='quench')
    assert LA.eval_struc_rate_anal(np.array(cat.variable_occs), 12) >= start


@pytest.mark.parametrize("variant", [0, 1, 4, 5])
def test_exact_guard_matches_reference_decisions(variant):
    """With the near-tie guard on, decisions follow the ORIGINAL fp64 weights even where the fixed-point delta is
    within the guard band of the threshold; the guard must fire and the result must equal the faithful oracle
    (every gated kernel: 0 = auto, 1 = row-per-lane, 4 = resident warp per chain, 5 = resident CTA per chain)."""
    from so_b200 import engine
    d, steps, n_chains = 12, 1440, 4
    z = np.load(os.path.join(Hp.GOLDEN, "sa_d12_golden.npz"))
    p = Hp.golden_surrogate(gated=True)
    lat = engine.Lattice(d, "LSC")
    ch = engine.Chains(lat, n_chains)
    ch.set_occupancy(z["x0"])
    ch.attach(Hp.tables_from_params(lat, p))
    stats, acc = ch.anneal(z["temps"], proposals=z["prop"], uniforms=z["u"], want_accept=True, exact_guard=True,
                           guard_tol=5e-3, variant=variant)           # wide band: the exact path takes many decisions
    assert stats["guard_fired"] > 10
    assert (acc.astype(bool) == z["gated_faithful_accept"]).all()
    assert (ch.get_occupancy() == z["gated_faithful_x"]).all()


def test_database_formats_roundtrip(tmp_path):
    """Files either side of the path (KMC_handler.py:13-55, Online_learn_main.py:26-47, generate_init_structures.py)."""
    from so_b200 import LSC_cat
    from so_b200.database import make_random_structure, write_structure_files, read_database
    cat = LSC_cat(12)
    folders = []
    for i in (4, 11):
        fldr = os.path.join(str(tmp_path), 'structure_%d' % i)
        x = make_random_structure(i, 48, cat, fldr)
        assert (np.array(x) == LA.make_random_structure(i, 48, 12)).all()          # same seeded recipe as upstream
        assert (np.load(os.path.join(fldr, 'occupancies.npy')) == np.array(x)).all()
        sym = np.load(os.path.join(fldr, 'occ_symmetries.npy'))
        assert sym.shape == (432, 144) and (sym == L.generate_all_translations_and_rotations(np.array(x), 12)).all()
        assert os.path.isfile(os.path.join(fldr, 'lattice_input.dat'))
        np.save(os.path.join(fldr, 'site_rates.npy'), cat.compute_site_rates_lsc())
        folders.append(fldr)
    os.makedirs(os.path.join(str(tmp_path), 'unfinished'))
    occs, rates = read_database(folders + [os.path.join(str(tmp_path), 'unfinished')])
    assert occs.shape == (864, 144) and rates.shape == (2, 144)


def test_lsc_analytical_batch_matches_oracle():
    """so_lsc_analytical (typing + dense steady-state solve per structure) vs OML/LSC_cat.py:260-316 restated."""
    from so_b200 import engine
    d = 12
    xs = np.array([LA.make_random_structure(i, 48, d) for i in range(0, 48, 3)] +
                  [np.zeros(144, np.uint8), np.ones(144, np.uint8)])
    ch = engine.Chains(engine.Lattice(d, "LSC"), len(xs))
    ch.set_occupancy(xs)
    rate, sr = ch.lsc_analytical()
    for c, x in enumerate(xs):
        ref = LA.compute_site_rates_lsc(x, d)
        np.testing.assert_allclose(sr[c], ref, rtol=1e-9, atol=1e-13)       # fp tolerance: 1e-9 relative
        np.testing.assert_allclose(rate[c], ref.sum(), rtol=1e-9, atol=1e-13)
    assert rate[-2] == 0.0 and rate[-1] == 0.0                              # no sites / terraces only: no edge flux


@pytest.mark.parametrize("d", [16, 24, 96])
def test_lsc_analytical_any_lattice_size(d):
    """OML/LSC_cat.py:260-316 works for any `dimen`; above 160 sites the device solves the same system by conjugate
    gradients on the lattice stencil (so_lsc.cu).  d = 16 against the reference-executed rates of ref_types.npz, all
    sizes against the LAPACK restatement, 1e-9 relative; then the SA loop on it and the drop-in LSC_cat(d)."""
    from so_b200 import engine, LSC_cat, optimize
    N = d * d
    rng = np.random.default_rng(d)
    n = 6 if d < 96 else 3
    xs = (rng.random((n, N)) < np.linspace(0.75, 0.97, n)[:, None]).astype(np.uint8)
    xs[0] = 1
    xs[0, [1, 2, d + 1]] = 0
    ch = engine.Chains(engine.Lattice(d, "LSC"), n)
    ch.set_occupancy(xs)
    rate, sr = ch.lsc_analytical()
    for c in range(n):
        ref = LA.compute_site_rates_lsc(xs[c], d)
        np.testing.assert_allclose(sr[c], ref, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(rate[c], ref.sum(), rtol=1e-9, atol=1e-12)
    if d == 16:
        z = np.load(os.path.join(Hp.GOLDEN, "ref_types.npz"))
        c2 = engine.Chains(engine.Lattice(d, "LSC"), len(z["d16_x"]))
        c2.set_occupancy(z["d16_x"])
        r2, s2 = c2.lsc_analytical()
        np.testing.assert_allclose(s2, z["d16_lsc_site_rates"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(r2, z["d16_lsc_struct_rate"], rtol=1e-9, atol=1e-12)
        cat = LSC_cat(16)                                                  # raised in round 1
        cat.assign_occs([int(v) for v in z["d16_x"][3]])
        cat.graph_to_KMClattice()
        np.testing.assert_allclose(cat.compute_site_rates_lsc(), z["d16_lsc_site_rates"][3], rtol=1e-9, atol=1e-12)
    if d <= 24:
        steps = 120
        prop = rng.integers(0, N, (n, steps)).astype(np.int32)
        u = rng.random((n, steps), dtype=np.float32)
        temps = SA.schedule("exp", 0.05, steps)
        stats, acc = ch.anneal_analytical(temps, proposals=prop, uniforms=u, want_accept=True)
        for c in range(0, n, 2):
            o = SA.optimize_analytical(xs[c], d, temps, prop[c], u[c])
            safe = o["margin"] > 1e-9
            assert (acc[c].astype(bool)[safe] == o["accept"][safe]).all() and safe.all()
            assert (ch.get_occupancy(c, 1)[0] == o["x"]).all()
            np.testing.assert_allclose(stats["OF"][c], o["OF"], rtol=1e-9, atol=1e-12)


def test_analytical_sa_matches_oracle():
    """optimize(mode='analytical') on the device: accept/reject sequence and final occupancy vs the literal loop."""
    from so_b200 import engine, LSC_cat, optimize
    d, steps, n = 12, 300, 3
    rng = np.random.default_rng(4)
    xs = np.array([LA.make_random_structure(i, 48, d) for i in (5, 17, 30)])
    prop = rng.integers(0, 144, (n, steps)).astype(np.int32)
    u = rng.random((n, steps), dtype=np.float32)
    temps = SA.schedule("exp", 0.05, steps)
    ch = engine.Chains(engine.Lattice(d, "LSC"), n)
    ch.set_occupancy(xs)
    stats, acc = ch.anneal_analytical(temps, proposals=prop, uniforms=u, want_accept=True)
    for c in range(n):
        o = SA.optimize_analytical(xs[c], d, temps, prop[c], u[c])
        safe = o["margin"] > 1e-9                                           # fp64 solves differ in the last bits
        assert (acc[c].astype(bool)[safe] == o["accept"][safe]).all() and safe.all()
        assert (ch.get_occupancy(c, 1)[0] == o["x"]).all()
        np.testing.assert_allclose(stats["OF"][c], o["OF"], rtol=1e-9)
    # the reference-facing call
    cat = LSC_cat(d)
    cat.assign_occs([int(v) for v in xs[0]])
    out = optimize(cat, mode='analytical', n_cycles=2, T_0=0.05, proposals=prop[0][:288], uniforms=u[0][:288],
                   n_record=4, return_stats=True)
    o = SA.optimize_analytical(xs[0], d, SA.schedule("exp", 0.05, 288), prop[0][:288], u[0][:288])
    assert cat.variable_occs == [int(v) for v in o["x"]] and abs(out["OF"] - o["OF"]) <= 1e-9 * abs(o["OF"])


def test_optimize_on_nh3_cat_default_lattice():
    """NH3_cat() default: 16x16, full 256-column translation descriptor (generic SA kernel, 16x16-window tensor-core
    state build); decisions and final structure against the C oracle, then the exported KMC site types."""
    from so_b200 import NH3_cat, optimize
    from oracle import c_oracle
    d, N = 16, 256
    p = S.SurrogateParams.random_init(L.full_offsets(d), H=20, seed=12, y_mean=0.1, y_scale=0.7)
    cat = NH3_cat()
    random.seed(2)
    cat.randomize(coverage=0.6)
    x0 = np.array(cat.variable_occs, dtype=np.uint8)
    n_cycles = 4
    steps = n_cycles * N
    rng = np.random.default_rng(3)
    prop = rng.integers(0, N, steps).astype(np.int32)
    u = rng.random(steps, dtype=np.float32)
    out = optimize(cat, surrogate=duck_surrogate(p), n_cycles=n_cycles, T_0=0.2, proposals=prop, uniforms=u,
                   exact_guard=False, return_stats=True)
    o = c_oracle.anneal_fixed_many(p, S.quantize(p), d, x0[None, :], SA.schedule("exp", 0.2, steps), prop[None, :], u[None, :])
    assert cat.variable_occs == [int(v) for v in o["x"][0]] and out["OF"] == o["OF"][0]
    assert out["accepted"] == int(o["accept"].sum())
    ref = S.eval_rate_fp64(p, L.generate_all_translations(o["x"][0], d))
    assert abs(out["OF"] - ref) <= 1e-5 * abs(ref)
    cat.graph_to_KMClattice()
    assert cat.KMC_lat.site_type_inds == ST.nh3_export(o["x"][0], d)


def test_active_learning_loop_runs():
    """The reference's learn -> anneal -> evaluate loop (drivers/Online_learn_main.py) through the OML module paths:
    database, device-fitted regressor + tree gate, SA on the GPU, analytical re-evaluation (scripts/active_learning_loop.py)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("al_loop", os.path.join(os.path.dirname(Hp.GOLDEN), "..", "scripts",
                                                                          "active_learning_loop.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    hist = mod.main(iterations=2, n_strucs=8, n_cycles=20, verbose=False)
    assert len(hist) == 2 and all(np.isfinite(v) for row in hist for v in row)
    assert hist[0][1] > hist[0][2]          # the first surrogate optimum beats every random island structure of the database
