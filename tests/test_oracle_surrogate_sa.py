"""CPU: surrogate forward and SA loop of the oracle against scikit-learn, the committed goldens and each other."""
import os
import numpy as np
import pytest
from oracle import lattice as L, surrogate as S, sa as SA, philox, lsc_analytical as LA
import helpers as Hp


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), "6627e8d5 e169c58d bc57ac4c 9b00dbd8"),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, "408f276d 41c83b0e a20bc7c6 6d5451fd"),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), "d16cfe09 94fdcceb 5001e420 24126ea1")]
    for c, k, want in kat:
        out = philox.philox4x32_10([np.array([v]) for v in c], [np.array([v]) for v in k])
        assert " ".join("%08x" % int(v[0]) for v in out) == want
    prop, u = philox.draws(7, np.arange(4)[:, None], np.arange(1000)[None, :], 144)
    assert prop.min() >= 0 and prop.max() < 144 and u.min() >= 0 and u.max() < 1 and u.dtype == np.float32
    assert abs(prop.mean() - 71.5) < 3 and abs(u.mean() - 0.5) < 0.02


def test_golden_surrogate_forward_and_sklearn_tree_semantics():
    p = Hp.golden_surrogate(gated=True)
    z = np.load(os.path.join(Hp.GOLDEN, "surrogate_d12.npz"))
    q = S.quantize(p)
    for x in z["train_occs"][:5]:
        T = L.generate_all_translations(x, 12)
        ref = S.eval_rate_fp64(p, T)
        of, hi, lo, n_act, gate = S.score_fixed(p, q, x, 12)
        assert abs(of - ref) <= 1e-5 * max(abs(ref), 1e-12)           # fixed-point arithmetic within the fp32 bar
        assert (gate == S.tree_predict(p.tree, T)).all() and n_act == gate.sum()
    # labels of the training set reproduce (LSC analytical oracle is deterministic)
    np.testing.assert_allclose(LA.compute_site_rates_lsc(z["train_occs"][7], 12), z["train_rates"][7], rtol=1e-12, atol=1e-15)
    assert (LA.make_random_structure(7, 48, 12) == z["train_occs"][7]).all()


def test_sklearn_predict_equals_extracted_arrays():
    occs = np.array([LA.make_random_structure(i, 48, 12) for i in range(0, 48, 8)])
    rates = np.array([LA.compute_site_rates_lsc(x, 12) for x in occs])
    syms = np.vstack([L.generate_all_translations_and_rotations(x, 12) for x in occs])
    sur = S.train(syms, rates, seed=0, max_iter=40)
    p = S.SurrogateParams.from_sklearn(sur.predictor, sur.Y_scaler, sur.classifier, L.full_offsets(12))
    for x in occs[:3]:
        T = L.generate_all_translations(x, 12)
        a, b = sur.eval_rate(T), S.eval_rate_fp64(p, T)
        assert abs(a - b) <= 1e-9 * max(1.0, abs(a))
        assert (S.tree_predict(p.tree, T) == (sur.classifier.predict(T) == 1.0)).all()


def test_quantisation_bounds():
    for K, H, seed in ((49, 20, 0), (144, 20, 1), (25, 7, 2)):
        off = L.local_offsets(3) if K == 49 else (L.local_offsets(2) if K == 25 else L.full_offsets(12))
        p = S.SurrogateParams.random_init(off, H=H, seed=seed)
        q = S.quantize(p)
        assert q["Hp"] % 4 == 0 and q["Hp"] >= H
        col = np.abs(q["b1q"].astype(np.int64)) + np.abs(q["W1q"].astype(np.int64)).sum(0)
        assert col.max() < 2 ** 31 and col.max() >= 2 ** 29            # uses the int32 range without overflow
        assert np.abs(q["w2q"]).max() <= 2 ** (31 - S.ceil_log2(float(q["Hp"])))
        np.testing.assert_allclose(np.ldexp(q["W1q"][:, :H].astype(float), q["e_h"]), p.W1, atol=2.0 ** (q["e_h"] - 1))
    assert [S.ceil_log2(v) for v in (1.0, 2.0, 3.0, 0.5, 0.75, 1e-3)] == [0, 1, 2, -1, 0, -9]


def test_sa_goldens_reproduce():
    z = np.load(os.path.join(Hp.GOLDEN, "sa_d12_golden.npz"))
    for name, gated in (("gated", True), ("open", False)):
        p = Hp.golden_surrogate(gated=gated)
        q = S.quantize(p)
        o = SA.anneal_fixed(p, q, 12, z["x0"][1], z["temps"], z["prop"][1], z["u"][1])
        assert (o["accept"] == z[name + "_fixed_accept"][1]).all() and (o["x"] == z[name + "_fixed_x"][1]).all()
        assert (o["hi"], o["lo"]) == (int(z[name + "_fixed_hi"][1]), int(z[name + "_fixed_lo"][1]))
        assert (z[name + "_fixed_accept"] == z[name + "_faithful_accept"]).all()       # fixed point == reference fp64
        np.testing.assert_allclose(z[name + "_fixed_OF"], z[name + "_faithful_OF"], rtol=1e-5)


def test_faithful_loop_short_run():
    """The literal restatement (O(N^2) copy + full evaluation) and the delta form take the same decisions."""
    p = Hp.golden_surrogate(gated=True)
    q = S.quantize(p)
    rng = np.random.default_rng(3)
    x0 = (rng.random(144) < 0.5).astype(np.uint8)
    steps = 150
    prop = rng.integers(0, 144, steps)
    u = rng.random(steps, dtype=np.float32)
    temps = SA.schedule("exp", 0.05, steps)

    class Sur(object):
        def eval_rate(self, M):
            return S.eval_rate_fp64(p, M)
    a = SA.optimize_faithful(x0, 12, Sur(), p.offsets, temps, prop, u, n_record=10)
    b = SA.anneal_fixed(p, q, 12, x0, temps, prop, u)
    assert (a["accept"] == b["accept"]).all() and (a["x"] == b["x"]).all()
    assert a["traj"].shape == (11, 3) and a["traj"][-1, 0] == steps
    # state invariants of the delta form
    of, hi, lo, n_act, _ = S.score_fixed(p, q, b["x"], 12)
    assert (hi, lo, n_act) == (b["hi"], b["lo"], b["n_act"]) and of == b["OF"]


def test_schedules():
    n = 1000
    np.testing.assert_allclose(SA.schedule("exp", 0.05, n)[-1], 0.05 * np.exp(-(n - 1) / (n / 5.0)))
    assert SA.schedule("log", 1.0, n)[0] == 1.0 / np.log(2) and (SA.schedule("quench", 1.0, n) == 0).all()
    lin = SA.schedule("linear", 2.0, n)
    assert lin[-1] == 0.0 and abs(lin[0] - 2.0 * (1 - 1.0 / n)) < 1e-15
    py2 = SA.schedule("linear", 2.0, n, py2_linear=True)
    assert (py2[:-1] == 2.0).all() and py2[-1] == 0.0                  # the upstream integer-division artefact
    with pytest.raises(NameError):
        SA.schedule("nope", 1.0, n)
    assert list(SA.steps_to_record(14400, 100)[:3]) == [0, 144, 288]
