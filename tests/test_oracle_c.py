"""CPU: the C restatement (oracle/so_oracle.c) against the NumPy oracle it mirrors and the committed goldens."""
import os
import numpy as np
from oracle import lattice as L, surrogate as S, sa as SA, c_oracle
import helpers as Hp


def test_c_oracle_equals_numpy_oracle_gated_and_open():
    z = np.load(os.path.join(Hp.GOLDEN, "sa_d12_golden.npz"))
    for name, gated in (("gated", True), ("open", False)):
        p = Hp.golden_surrogate(gated=gated)
        q = S.quantize(p)
        o = c_oracle.anneal_fixed_many(p, q, 12, z["x0"], z["temps"], z["prop"], z["u"])
        assert (o["accept"] == z[name + "_fixed_accept"]).all() and (o["x"] == z[name + "_fixed_x"]).all()
        assert (o["hi"] == z[name + "_fixed_hi"]).all() and (o["lo"] == z[name + "_fixed_lo"]).all()
        assert (o["OF"] == z[name + "_fixed_OF"]).all()


def test_c_oracle_local_descriptor_with_tscale():
    d = 9
    p = S.SurrogateParams.random_init(L.local_offsets(3), H=7, seed=3, y_mean=0.2, y_scale=1.3)
    q = S.quantize(p)
    rng = np.random.default_rng(0)
    x0 = (rng.random((5, d * d)) < 0.5).astype(np.uint8)
    steps = 400
    prop = rng.integers(0, d * d, (5, steps)).astype(np.int32)
    u = rng.random((5, steps), dtype=np.float32)
    temps = SA.schedule("exp", 0.5, steps)
    ts = np.linspace(0.3, 2.0, 5)
    o = c_oracle.anneal_fixed_many(p, q, d, x0, temps, prop, u, tscale=ts)
    for c in range(5):
        r = SA.anneal_fixed(p, q, d, x0[c], temps, prop[c], u[c], tscale=ts[c])
        assert (o["accept"][c] == r["accept"]).all() and (o["x"][c] == r["x"]).all()
        assert (int(o["hi"][c]), int(o["lo"][c]), int(o["n_act"][c])) == (r["hi"], r["lo"], r["n_act"]) and o["OF"][c] == r["OF"]


def test_c_fp64_rule_equals_reference_executed_optimize():
    """oracle_anneal_fp64 (the reference's fp64 rule with its original weights, the decision oracle of the full-size GPU
    test) against optimize() EXECUTED FROM THE REFERENCE SOURCE (tests/golden/ref_surrogate_sa.npz): every run."""
    from test_ref_golden import ref_surrogate, _sa_runs
    z = np.load(os.path.join(Hp.GOLDEN, "ref_surrogate_sa.npz"))
    worst = np.inf
    for name, c, gated, n_cycles, T_0, cooling in _sa_runs(z):
        steps = n_cycles * 144
        py2 = cooling == "linear_py2"
        temps = SA.schedule("linear" if py2 else cooling, T_0, steps, py2_linear=py2)
        p = ref_surrogate(z, gated)
        o = c_oracle.anneal_fp64_many(p, 12, z["sa_x0"][c:c + 1], temps, z["sa_prop"][c:c + 1, :steps], z["sa_u"][c:c + 1, :steps])
        acc = np.unpackbits(z["sa_%s_accept" % name])[:steps].astype(bool)
        assert (o["accept"][0] == acc).all() and (o["x"][0] == z["sa_%s_x" % name]).all(), name
        assert o["n_below"][0] == 0
        worst = min(worst, float(o["min_margin"][0]))
        of_new = z["sa_%s_OF_new" % name]
        cur = float(z["sa_%s_OF0" % name])
        for s in range(steps):
            if acc[s]:
                cur = of_new[s]
        assert abs(o["OF"][0] - cur) <= 1e-9 * max(1.0, abs(cur))
    assert worst > 1e-7
