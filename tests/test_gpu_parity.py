"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle on the same seeded inputs.
Integer/byte/index work must be bit-exact; scores within 1e-5 relative of the fp64 reference arithmetic."""
import os
import numpy as np
import pytest

from oracle import lattice as L, site_types as ST, surrogate as S, sa as SA, philox
import helpers as Hp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from so_b200 import engine
    return engine


# ---- lattice ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d", [4, 5, 12])
def test_lattice_tables(eng, d):
    lat = eng.Lattice(d, "NH3")
    rowptr, col = lat.csr()
    rp, cl = L.neighbour_csr(d)
    assert (rowptr == rp).all()
    for i in range(5 * d * d):
        assert set(col[rowptr[i]:rowptr[i + 1]].tolist()) == set(cl[rp[i]:rp[i + 1]].tolist())
    pos, cell = lat.positions()
    np.testing.assert_allclose(pos, L.slab_positions(d), rtol=0, atol=1e-12)
    np.testing.assert_allclose(cell, L.cell_vectors(d), rtol=0, atol=1e-12)


def test_lattice_rejects_tiny(eng):
    with pytest.raises(ValueError):
        eng.Lattice(3, "LSC")


# ---- occupancy -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d", [5, 12, 16])
def test_occupancy_roundtrip(eng, d):
    lat = eng.Lattice(d, "LSC")
    n = 37
    ch = eng.Chains(lat, n)
    assert (ch.get_occupancy() == 1).all()                      # all atoms present (LSC_cat.py:50)
    x = Hp.random_structures(n, d * d, seed=d)
    ch.set_occupancy(x)
    assert (ch.get_occupancy() == x).all()
    assert (ch.get_bits() == eng.pack_bits(x)).all()
    ch.set_bits(eng.pack_bits(x[::-1].copy()))
    assert (ch.get_occupancy() == x[::-1]).all()
    ch.set_occupancy(x[3:5], chain0=10)                         # ragged sub-range
    assert (ch.get_occupancy(10, 2) == x[3:5]).all()
    bad = x.copy()
    bad[2, 7] = 2
    with pytest.raises(NameError):
        ch.set_occupancy(bad)
    with pytest.raises(ValueError):
        ch.set_occupancy(x, chain0=1)                           # runs past the end


def test_flip(eng):
    d = 12
    lat = eng.Lattice(d, "LSC")
    ch = eng.Chains(lat, 4)
    x = Hp.random_structures(4, 144, seed=1)
    ch.set_occupancy(x)
    ch.flip([0, 0, 3, 2, 0], [5, 6, 143, 0, 5])                  # site 5 of chain 0 toggles twice
    y = x.copy()
    for c, s in zip([0, 0, 3, 2, 0], [5, 6, 143, 0, 5]):
        y[c, s] ^= 1
    assert (ch.get_occupancy() == y).all()
    with pytest.raises(ValueError):
        ch.flip([0], [144])


def test_randomize_statistics(eng):
    lat = eng.Lattice(12, "LSC")
    ch = eng.Chains(lat, 2048)
    ch.randomize(seed=3, coverage=0.5)
    x = ch.get_occupancy()
    assert abs(x.mean() - 0.5) < 0.01
    ch.randomize(seed=3, coverage=0.5)
    assert (ch.get_occupancy() == x).all()                      # counter-based: same seed, same bits
    ch.randomize(seed=4, coverage=None)
    cov = ch.get_occupancy().mean(1)
    assert cov.std() > 0.2 and 0.4 < cov.mean() < 0.6           # per-chain coverage ~ U(0,1)


# ---- descriptors / site types (bit-exact) -------------------------------------------------------------------
@pytest.mark.parametrize("d", [4, 7, 12, 16, 33])
def test_descriptors_bit_exact(eng, d):
    N = d * d
    lat = eng.Lattice(d, "NH3")
    x = np.vstack([Hp.random_structures(60, N, seed=100 + d), np.zeros((1, N), np.uint8), np.ones((1, N), np.uint8)])
    ch = eng.Chains(lat, len(x))
    ch.set_occupancy(x)
    out = ch.descriptors(c6=True, lsc_type=True, nh3_type=True, lsc_hist=True, nh3_hist=True, nh3_exp=True)
    assert (out["c6"] == ST.c6_counts(x, d)).all()
    lsc = ST.lsc_types(x, d)
    assert (out["lsc_type"] == lsc).all()
    nh3 = ST.nh3_types(x, d)
    assert (out["nh3_type"] == nh3).all()
    assert (out["lsc_hist"] == np.stack([(lsc == t).sum(1) for t in range(3)], 1)).all()
    assert (out["nh3_hist"] == ST.nh3_hist(nh3)).all()
    assert (out["nh3_exp"] == ST.nh3_export_hist(nh3)).all()


def test_descriptors_golden(eng):
    z = np.load(os.path.join(Hp.GOLDEN, "types_golden.npz"))
    for d in (6, 8):
        x = z["x_d%d" % d]
        ch = eng.Chains(eng.Lattice(d, "NH3"), len(x))
        ch.set_occupancy(x)
        out = ch.descriptors(lsc_type=True, nh3_type=True)
        assert (out["lsc_type"] == z["lsc_d%d" % d]).all()
        assert (out["nh3_type"] == z["nh3_d%d" % d]).all()


@pytest.mark.parametrize("d", [4, 5, 12, 16])
def test_flip_delta_bit_exact(eng, d):
    N = d * d
    n = 300
    lat = eng.Lattice(d, "NH3")
    x = Hp.random_structures(n, N, seed=7 + d)
    sites = np.random.default_rng(d).integers(0, N, n).astype(np.int32)
    ch = eng.Chains(lat, n)
    ch.set_occupancy(x)
    d_lsc, d_nh3 = ch.flip_delta(sites)
    y = x.copy()
    y[np.arange(n), sites] ^= 1
    lsc0, lsc1 = ST.lsc_types(x, d), ST.lsc_types(y, d)
    exp_lsc = np.stack([(lsc1 == t).sum(1) - (lsc0 == t).sum(1) for t in range(3)], 1)
    e0, e1 = ST.nh3_export_hist(ST.nh3_types(x, d)), ST.nh3_export_hist(ST.nh3_types(y, d))
    assert (d_lsc == exp_lsc).all()
    assert (d_nh3 == e1 - e0).all()
    assert (ch.get_occupancy() == x).all()                       # evaluation only, nothing applied


# ---- surrogate scores -------------------------------------------------------------------------------------
def _check_scores(eng, d, p, x):
    lat = eng.Lattice(d, "LSC")
    tab = Hp.tables_from_params(lat, p)
    q = S.quantize(p)
    assert (tab.e_h, tab.e_w, tab.Hp) == (q["e_h"], q["e_w"], q["Hp"])
    ch = eng.Chains(lat, len(x))
    ch.set_occupancy(x)
    ch.attach(tab)
    of, hi, lo, na = ch.objective(sums=True)
    for c in range(len(x)):
        o_of, o_hi, o_lo, o_n, _ = S.score_fixed(p, q, x[c], d)
        assert (int(hi[c]), int(lo[c]), int(na[c])) == (o_hi, o_lo, o_n)       # exact integer sums
        assert of[c] == o_of                                                   # pinned fp64 conversion
        ref = S.eval_rate_fp64(p, L.descriptor_rows(x[c], p.offsets, d))
        assert abs(of[c] - ref) <= 1e-5 * max(abs(ref), 1e-12) or (ref == 0 and of[c] == 0)
    hq, _ = S.preact_fixed(p, q, x[0], d)
    assert (ch.preact(0) == hq).all()
    assert (ch.score() == of).all()                                            # recompute == tracked
    rate, lh, ne = tab.score_batch(eng.pack_bits(x), want_lsc_hist=True)
    assert (rate == of).all()
    lsc = ST.lsc_types(x, d)
    assert (lh == np.stack([(lsc == t).sum(1) for t in range(3)], 1)).all()
    return ch, tab


def test_scores_full_descriptor_gated(eng):
    p = Hp.golden_surrogate(gated=True)
    x = np.vstack([Hp.random_structures(40, 144, seed=11), np.zeros((1, 144), np.uint8), np.ones((1, 144), np.uint8)])
    _check_scores(eng, 12, p, x)


def test_scores_full_descriptor_open(eng):
    _check_scores(eng, 12, Hp.golden_surrogate(gated=False), Hp.random_structures(40, 144, seed=12))


@pytest.mark.parametrize("d,H", [(12, 20), (16, 20), (9, 7), (24, 32)])
def test_scores_local_descriptor(eng, d, H):
    p = S.SurrogateParams.random_init(L.local_offsets(3), H=H, seed=d, y_mean=0.3, y_scale=1.7)
    _check_scores(eng, d, p, Hp.random_structures(24, d * d, seed=13 + d))


def test_eval_rows(eng):
    p = Hp.golden_surrogate(gated=True)
    lat = eng.Lattice(12, "LSC")
    tab = Hp.tables_from_params(lat, p)
    X = (np.random.default_rng(0).random((500, 144)) < 0.6).astype(np.uint8)
    active, rate = tab.eval_rows(X)
    q = S.quantize(p)
    assert (active == S.tree_predict(p.tree, X)).all()
    ref = np.maximum(X @ p.W1 + p.b1, 0) @ p.w2 + p.b2
    np.testing.assert_allclose(rate, ref * p.y_scale + p.y_mean, rtol=1e-5, atol=1e-7)


# ---- annealing: decisions, occupancies, sums bit-exact against the fixed-point oracle --------------------------
def _run_sa(eng, d, p, n_chains, steps, seed, T0=0.05, variant=0, use_philox=False, exact_guard=False,
            check_chains=None, cooling="exp"):
    N = d * d
    rng = np.random.Generator(np.random.Philox(key=seed))
    x0 = (rng.random((n_chains, N)) < 0.5).astype(np.uint8)
    temps = SA.schedule(cooling, T0, steps)
    if use_philox:
        prop, u = philox.draws(seed, np.arange(n_chains)[:, None] + 1000, np.arange(steps)[None, :] + 5, N)
    else:
        prop = rng.integers(0, N, size=(n_chains, steps)).astype(np.int32)
        u = rng.random((n_chains, steps), dtype=np.float32)
    lat = eng.Lattice(d, "LSC")
    tab = Hp.tables_from_params(lat, p)
    ch = eng.Chains(lat, n_chains)
    ch.set_occupancy(x0)
    ch.attach(tab)
    if use_philox:
        stats, acc = ch.anneal(temps, step0=5, seed=seed, chain_id0=1000, want_accept=True, variant=variant)
    else:
        stats, acc = ch.anneal(temps, proposals=prop, uniforms=u, want_accept=True, variant=variant,
                               exact_guard=exact_guard, guard_tol=1e-6)
    q = S.quantize(p)
    x1 = ch.get_occupancy()
    of, hi, lo, na = ch.objective(sums=True)
    n_acc = int(acc.sum())
    for c in (range(n_chains) if check_chains is None else check_chains):
        o = SA.anneal_fixed(p, q, d, x0[c], temps, prop[c], u[c])
        if not exact_guard:
            assert (acc[c].astype(bool) == o["accept"]).all(), "chain %d: accept/reject sequence differs" % c
            assert (x1[c] == o["x"]).all()
            assert (int(hi[c]), int(lo[c]), int(na[c])) == (o["hi"], o["lo"], o["n_act"])
            assert of[c] == o["OF"]
            if c == 0:
                assert (ch.preact(0) == o["hq"]).all()
    assert stats["accepted"] == n_acc and stats["flips"] == n_chains * steps
    # state invariant: tracked state == fresh recompute from the final bits
    tracked = ch.objective(sums=True)
    fresh_of = ch.score()
    again = ch.objective(sums=True)
    assert (tracked[0] == fresh_of).all() and all((a == b).all() for a, b in zip(tracked, again))
    return dict(x0=x0, x1=x1, acc=acc, prop=prop, u=u, temps=temps, of=of, stats=stats)


@pytest.mark.parametrize("variant", [0, 2])
def test_sa_flat_full_descriptor(eng, variant):
    _run_sa(eng, 12, Hp.golden_surrogate(gated=False), n_chains=24, steps=1500, seed=21, variant=variant)


def test_sa_flat_local_descriptor(eng):
    p = S.SurrogateParams.random_init(L.local_offsets(3), H=20, seed=3)
    _run_sa(eng, 16, p, n_chains=40, steps=2000, seed=22)


# variant 0 picks a shared-memory-resident kernel on these small lattices (CTA per chain for few chains, warp per chain
# for many), variant 4 the warp-per-chain one, variant 3 the pipelined window kernel
@pytest.mark.parametrize("variant", [0, 3, 4])
@pytest.mark.parametrize("T0", [0.05, 2.0])
def test_sa_flat_local_small_lattice_wrap(eng, T0, variant):
    """7x7 window on a 7x7 lattice: every wrap path, and (hot run: most flips accepted) every prefetched window is
    stale -- the hazard paths of the pipelined kernel."""
    p = S.SurrogateParams.random_init(L.local_offsets(3), H=20, seed=4)
    r = _run_sa(eng, 7, p, n_chains=16, steps=800, seed=23, T0=T0, cooling="const" if T0 > 1 else "exp", variant=variant)
    if T0 > 1:
        assert r["acc"].mean() > 0.5


@pytest.mark.parametrize("variant", [0, 3, 4])
@pytest.mark.parametrize("d", [8, 12, 20])
def test_sa_window_hot_chains(eng, d, variant):
    """High acceptance on small lattices: overlapping windows of consecutive flips are the common case."""
    p = S.SurrogateParams.random_init(L.local_offsets(3), H=20, seed=40 + d)
    r = _run_sa(eng, d, p, n_chains=32, steps=1500, seed=50 + d, T0=1.0, cooling="const", variant=variant)
    assert r["acc"].mean() > 0.1


@pytest.mark.parametrize("variant,steps", [(0, 70), (3, 70), (3, 71), (3, 72), (4, 70)])
def test_sa_window_many_chains(eng, variant, steps):
    """More chains than resident warps: dynamic chain hand-out + stage/mbarrier reuse across chains (the ring of stages is
    continuous over the chains of a warp: all three step counts modulo the number of stages)."""
    p = S.SurrogateParams.random_init(L.local_offsets(3), H=20, seed=41)
    _run_sa(eng, 12, p, n_chains=5000, steps=steps, seed=51, T0=0.5, check_chains=range(0, 5000, 97), cooling="const",
            variant=variant)


@pytest.mark.parametrize("variant", [0, 1])
def test_sa_gated_full_descriptor(eng, variant):
    _run_sa(eng, 12, Hp.golden_surrogate(gated=True), n_chains=16, steps=1200, seed=24, variant=variant)


def _random_tree(K, depth, seed, grow=0.8):
    rng = np.random.default_rng(seed)
    feature, thr, left, right, active = [], [], [], [], []

    def node(level):
        i = len(feature)
        feature.append(-2); thr.append(0.0); left.append(-1); right.append(-1); active.append(int(rng.random() < 0.6))
        # uneven depth, plus a few tests that ignore their bit (thr outside [0, 1))
        if level < depth and (level < 2 or rng.random() < grow):
            feature[i] = int(rng.integers(0, K))
            thr[i] = float(rng.choice([0.5, 0.5, 0.5, 0.5, -1.0, 1.5]))
            left[i] = node(level + 1)
            right[i] = node(level + 1)
        return i
    node(0)
    return dict(feature=np.array(feature, np.int32), thr=np.array(thr), left=np.array(left, np.int32),
                right=np.array(right, np.int32), active=np.array(active, np.uint8))


@pytest.mark.parametrize("variant", [0, 1, 3, 6])
@pytest.mark.parametrize("d,T0", [(24, 0.05), (24, 2.0), (40, 0.3), (48, 1.0)])
def test_sa_gated_local_large_lattice(eng, d, T0, variant):
    """Tree gate on lattices too large for the warp-per-chain resident kernel: 20 chains run on the CTA-per-chain
    kernel (variant 0, state in up to 227 KB of shared memory), the gated window kernel (variant 3: gate records in the
    TMA pipeline), the cached-gate kernel with global state (variant 6: gate bits + path index, only active rows read)
    and the row-per-lane kernel (variant 1), all against the oracle; hot runs exercise the record / index upkeep on accept."""
    p = S.SurrogateParams.random_init(L.local_offsets(3), H=20, seed=60 + d)
    p.tree = _random_tree(49, 7, seed=d)
    r = _run_sa(eng, d, p, n_chains=20, steps=1500, seed=70 + d, T0=T0, variant=variant,
                cooling="const" if T0 > 1 else "exp")
    if T0 > 1:
        assert r["acc"].mean() > 0.1


@pytest.mark.parametrize("compact", [False, True])
@pytest.mark.parametrize("d,n,steps,T0,depth", [(7, 24, 900, 2.0, 8), (9, 24, 900, 1.0, 6), (13, 40, 900, 2.0, 9), (20, 300, 500, 1.0, 8),
                                                (33, 600, 300, 0.5, 8), (96, 48, 1500, 1.0, 8)])
def test_sa_window_gated_kernel(eng, d, n, steps, T0, depth, compact):
    """The gated window kernel (tree gate + 7x7 window, state and per-row gate records through the TMA pipeline) on the
    int32 and on the 24-bit packed state: accept vectors, occupancies, exact sums, active-row counts and the whole
    first-layer state equal the fixed-point oracle; hot chains on lattices from 7x7 (every window wraps, the bit patch
    holds the flipped site several times, every prefetched window is stale) to 96x96; then the same steps in several
    launches (records kept between launches), after outside changes of the bits (records rebuilt), and from Philox draws."""
    from oracle import c_oracle
    N = d * d
    p = S.SurrogateParams.random_init(L.local_offsets(3), H=20, seed=300 + d, y_mean=0.05, y_scale=1.3)
    p.tree = _random_tree(49, depth, seed=400 + d)
    q = S.quantize(p, bits="packed24") if compact else S.quantize(p)
    lat = eng.Lattice(d, "NH3")
    tab = eng.SurrogateTables(lat, p.offsets, p.W1, p.b1, p.w2, p.b2, p.y_mean, p.y_scale, tree=p.tree, compact=compact)
    rng = np.random.default_rng(d)
    x0 = (rng.random((n, N)) < 0.5).astype(np.uint8)
    prop = rng.integers(0, N, (n, steps)).astype(np.int32)
    u = rng.random((n, steps), dtype=np.float32)
    temps = SA.schedule("const", T0, steps)
    ch = eng.Chains(lat, n)
    ch.set_occupancy(x0)
    ch.attach(tab)
    ch.track_types(True)
    st, acc = ch.anneal(temps, proposals=prop, uniforms=u, want_accept=True, variant=3)
    assert st["kernel"] == "sa_window_gated_kernel" and st["accepted"] > 0.05 * st["flips"]
    o = c_oracle.anneal_fixed_many(p, q, d, x0, temps, prop, u)
    assert (acc.astype(bool) == o["accept"]).all()
    x1 = ch.get_occupancy()
    assert (x1 == o["x"]).all()
    of, hi, lo, na = ch.objective(sums=True)
    assert (hi == o["hi"]).all() and (lo == o["lo"]).all() and (of == o["OF"]).all() and (na == o["n_act"]).all()
    hq1, _ = S.preact_fixed(p, q, x1[n // 2], d)
    assert (ch.preact(n // 2)[:, :20] == hq1[:, :20]).all()
    assert (ch.score() == of).all()
    lh, ne = ch.type_histograms()
    assert (ne == ST.nh3_export_hist(ST.nh3_types(x1, d))).all()
    # the same steps in four launches: the records stay valid between launches of this kernel ...
    ch.set_occupancy(x0)
    cuts = (0, steps // 3, steps // 3 + 1, steps // 2, steps)
    for a, b in zip(cuts[:-1], cuts[1:]):
        s2 = ch.anneal(temps[a:b], step0=a, proposals=prop[:, a:b], uniforms=u[:, a:b], variant=3)
        assert s2["kernel"] == "sa_window_gated_kernel"
    assert (ch.get_occupancy() == x1).all() and (ch.objective() == of).all()
    # ... and are rebuilt after the bits changed behind its back (another kernel, set_occupancy)
    if not compact:
        ch.set_occupancy(x0)
        h = steps // 2
        ch.anneal(temps[:h], proposals=prop[:, :h], uniforms=u[:, :h], variant=1)
        ch.anneal(temps[h:], step0=h, proposals=prop[:, h:], uniforms=u[:, h:], variant=3)
        assert (ch.get_occupancy() == x1).all() and (ch.objective() == of).all()
    ch.set_occupancy(x0)
    ch.anneal(temps, seed=3, variant=3)
    xa, ofa = ch.get_occupancy(), ch.objective()
    ch.set_occupancy(x0)
    ch.anneal(temps, seed=3, variant=3)
    assert (ch.get_occupancy() == xa).all() and (ch.objective() == ofa).all()
    assert (ch.score() == ofa).all()
    # near ties re-decided from the fp64 weights: decisions equal the fp64 rule
    ch.set_occupancy(x0)
    tol = 8.0 * np.sqrt(tab.K) * np.ldexp(1.0, tab.e_h) * abs(p.y_scale) * float(np.abs(p.w2).sum())
    m = min(n, 16)
    _, acc2 = ch.anneal(temps, proposals=prop, uniforms=u, want_accept=True, exact_guard=True, guard_tol=tol, variant=3)
    o64 = c_oracle.anneal_fp64_many(p, d, x0[:m], temps, prop[:m], u[:m], margin_thr=1e-9)
    ok = o64["n_below"] == 0
    assert (acc2[:m].astype(bool)[ok] == o64["accept"][ok]).all() and ok.mean() > 0.9


def test_sa_window_gated_many_chains(eng):
    """More chains than resident warps on the gated window kernel: dynamic chain hand-out, the two-stage ring and its mbarrier
    phases carried from one chain of a warp to the next, records rebuilt per chain."""
    p = S.SurrogateParams.random_init(L.local_offsets(3), H=20, seed=141)
    p.tree = _random_tree(49, 7, seed=142)
    for steps in (70, 71):                       # even and odd step counts: both ring positions at the hand-over
        r = _run_sa(eng, 12, p, n_chains=6000, steps=steps, seed=151 + steps, T0=0.5, check_chains=range(0, 6000, 113),
                    cooling="const", variant=3)
        assert r["stats"]["kernel"] == "sa_window_gated_kernel"


@pytest.mark.parametrize("variant", [4, 5])
@pytest.mark.parametrize("d", [8, 12])
def test_sa_gated_hot_resident(eng, d, variant):
    """Resident gated kernels (4: warp per chain, 5: CTA per chain) at high acceptance (index upkeep every few steps),
    local window with wraps."""
    p = S.SurrogateParams.random_init(L.local_offsets(3), H=20, seed=80 + d)
    p.tree = _random_tree(49, 8, seed=100 + d)
    r = _run_sa(eng, d, p, n_chains=24, steps=1500, seed=90 + d, T0=2.0, cooling="const", variant=variant)
    assert r["acc"].mean() > 0.1


@pytest.mark.parametrize("d,variant", [(12, 4), (12, 5), (24, 0), (24, 3)])
def test_sa_gated_large_tree(eng, d, variant):
    """A tree too large for shared memory (8,191 nodes > 2,048) is walked in global memory by the same kernels."""
    p = S.SurrogateParams.random_init(L.local_offsets(3), H=20, seed=120 + d)
    p.tree = _random_tree(49, 12, seed=121, grow=1.0)
    assert len(p.tree["feature"]) == 8191
    r = _run_sa(eng, d, p, n_chains=12, steps=1200, seed=122 + d, T0=1.0, cooling="const", variant=variant)
    assert r["acc"].mean() > 0.1


@pytest.mark.parametrize("variant", [4, 5])
@pytest.mark.parametrize("gated", [False, True])
def test_sa_resident_full_descriptor_variants(eng, gated, variant):
    """Full 144-column descriptor (the reference's shape), golden MLP (+ tree), both resident kernels, Philox draws and
    injected draws, odd lattice for ragged row groups."""
    _run_sa(eng, 12, Hp.golden_surrogate(gated=gated), n_chains=10, steps=1200, seed=31, variant=variant, T0=0.2)
    _run_sa(eng, 12, Hp.golden_surrogate(gated=gated), n_chains=5, steps=700, seed=32, variant=variant, use_philox=True)
    p = S.SurrogateParams.random_init(L.local_offsets(2), H=7, seed=33)
    if gated:
        p.tree = _random_tree(25, 5, seed=34)
    _run_sa(eng, 9, p, n_chains=7, steps=900, seed=35, variant=variant, T0=1.0, cooling="const")


def test_sa_kernel_variants_agree(eng):
    """CTA-per-chain (auto choice for 12 chains) == resident warp-per-chain == window (pipelined) == generic flat ==
    row-per-lane kernels, bit for bit."""
    p = S.SurrogateParams.random_init(L.local_offsets(3), H=20, seed=5)
    a = _run_sa(eng, 12, p, n_chains=12, steps=1000, seed=25, variant=0, T0=0.3)
    for variant in (1, 2, 3, 4, 5):
        b = _run_sa(eng, 12, p, n_chains=12, steps=1000, seed=25, variant=variant, T0=0.3)
        assert (a["acc"] == b["acc"]).all() and (a["x1"] == b["x1"]).all() and (a["of"] == b["of"]).all()


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4])
def test_sa_philox_draws(eng, variant):
    p = S.SurrogateParams.random_init(L.local_offsets(3), H=20, seed=6)
    _run_sa(eng, 12, p, n_chains=12, steps=1000, seed=26, use_philox=True, variant=variant)


def test_sa_odd_hidden_width(eng):
    p = S.SurrogateParams.random_init(L.local_offsets(2), H=7, seed=7)
    _run_sa(eng, 10, p, n_chains=8, steps=600, seed=27)


def test_sa_quench_and_ties(eng):
    """T == 0: accept iff delta > 0 (optimize_SA.py:113-116)."""
    d, p = 12, Hp.golden_surrogate(gated=True)
    rng = np.random.default_rng(9)
    x0 = Hp.random_structures(6, 144, seed=9, coverage=0.5)
    steps = 400
    prop = rng.integers(0, 144, (6, steps)).astype(np.int32)
    u = rng.random((6, steps), dtype=np.float32)
    temps = SA.schedule("quench", 0.05, steps)
    lat = eng.Lattice(d, "LSC")
    ch = eng.Chains(lat, 6)
    ch.set_occupancy(x0)
    ch.attach(Hp.tables_from_params(lat, p))
    of0 = ch.objective()
    stats, acc = ch.anneal(temps, proposals=prop, uniforms=u, want_accept=True)
    q = S.quantize(p)
    for c in range(6):
        o = SA.anneal_fixed(p, q, d, x0[c], temps, prop[c], u[c])
        assert (acc[c].astype(bool) == o["accept"]).all()
    assert (ch.objective() >= of0).all()


def test_sa_golden_vectors(eng):
    """Committed oracle trajectories (faithful fp64 == fixed point) reproduced by the GPU."""
    z = np.load(os.path.join(Hp.GOLDEN, "sa_d12_golden.npz"))
    d = 12
    for name, gated in (("gated", True), ("open", False)):
        p = Hp.golden_surrogate(gated=gated)
        lat = eng.Lattice(d, "LSC")
        ch = eng.Chains(lat, len(z["x0"]))
        ch.set_occupancy(z["x0"])
        ch.attach(Hp.tables_from_params(lat, p))
        stats, acc = ch.anneal(z["temps"], proposals=z["prop"], uniforms=z["u"], want_accept=True)
        assert (acc.astype(bool) == z[name + "_fixed_accept"]).all()
        assert (acc.astype(bool) == z[name + "_faithful_accept"]).all()
        assert (ch.get_occupancy() == z[name + "_faithful_x"]).all()
        of, hi, lo, _ = ch.objective(sums=True)
        assert (hi == z[name + "_fixed_hi"]).all() and (lo == z[name + "_fixed_lo"]).all()
        np.testing.assert_allclose(of, z[name + "_faithful_OF"], rtol=1e-5)


def test_tscale_and_best(eng):
    d = 12
    p = S.SurrogateParams.random_init(L.local_offsets(3), H=20, seed=8)
    lat = eng.Lattice(d, "LSC")
    ch = eng.Chains(lat, 50)
    x = Hp.random_structures(50, 144, seed=30)
    ch.set_occupancy(x)
    ch.attach(Hp.tables_from_params(lat, p))
    of = ch.objective()
    best_of, best_chain, bits = ch.best()
    assert best_chain == int(np.argmax(of)) and best_of == of.max()
    assert (eng.unpack_bits(bits[None, :], 144)[0] == x[best_chain]).all()
    ts = np.linspace(0.5, 2.0, 50)
    ch.set_tscale(ts)
    assert (ch.get_tscale() == ts).all()
    # tscale scales the temperature: T = tscale * temps
    steps = 300
    rng = np.random.default_rng(2)
    prop = rng.integers(0, 144, (50, steps)).astype(np.int32)
    u = rng.random((50, steps), dtype=np.float32)
    temps = SA.schedule("const", 0.05, steps)
    stats, acc = ch.anneal(temps, proposals=prop, uniforms=u, want_accept=True)
    q = S.quantize(p)
    for c in (0, 17, 49):
        o = SA.anneal_fixed(p, q, d, x[c], temps, prop[c], u[c], tscale=ts[c])
        assert (acc[c].astype(bool) == o["accept"]).all()


# ---- replica exchange / best structure (new capability; oracle/exchange.py is the checker) ---------------------
@pytest.mark.parametrize("world,rank", [(1, 0), (2, 1), (4, 2)])
def test_exchange_round_matches_oracle(eng, world, rank):
    """One process emulates `rank` of `world`: the all-gathered arrays are fabricated on the device, the kernel must
    take the oracle's swap decisions and update exactly this rank's chains."""
    import ctypes as C
    import torch
    from oracle import exchange as X
    from so_b200 import _lib as B
    d, n_local, ladder = 12, 48, 8
    lat = eng.Lattice(d, "LSC")
    ch = eng.Chains(lat, n_local)
    p = S.SurrogateParams.random_init(L.local_offsets(3), H=20, seed=9)
    ch.attach(Hp.tables_from_params(lat, p))
    rng = np.random.default_rng(world * 10 + rank)
    of_all = rng.normal(size=world * n_local)
    g = np.concatenate([np.arange(n_local) * world + r for r in range(world)])
    ts = np.exp(-5.0 * (g % ladder) / (ladder - 1.0))
    rng.shuffle(ts.reshape(world * n_local // ladder, ladder).T)              # scramble rung order inside ladders
    ch.set_tscale(ts[rank * n_local:(rank + 1) * n_local])
    of_dev, ts_dev = torch.from_numpy(of_all).cuda(), torch.from_numpy(ts.copy()).cuda()
    t_ref, seed = 0.05, 77
    swapped_total = 0
    cur = ts.copy()
    for rnd in range(4):
        want = X.exchange_round(of_all, cur, world, n_local, ladder, rnd & 1, seed, rnd, t_ref)
        ns = C.c_int64()
        B.check(ch.lib.so_exchange(ch.h, rank, world, ladder, rnd & 1, seed, rnd, t_ref, C.c_void_p(of_dev.data_ptr()),
                                   C.c_void_p(ts_dev.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream),
                                   C.byref(ns)))
        assert (ts_dev.cpu().numpy() == want).all()
        assert (ch.get_tscale() == want[rank * n_local:(rank + 1) * n_local]).all()
        assert ns.value == int((want != cur)[rank * n_local:(rank + 1) * n_local].sum())
        swapped_total += ns.value
        cur = want
    assert swapped_total > 0 and sorted(cur.tolist()) == sorted(ts.tolist())          # temperatures are only permuted


def test_best_record_and_objectives_dev(eng):
    import ctypes as C
    import torch
    from so_b200 import parallel as P, _lib as B
    d, n = 12, 300
    lat = eng.Lattice(d, "LSC")
    ch = eng.Chains(lat, n)
    x = Hp.random_structures(n, 144, seed=5)
    ch.set_occupancy(x)
    p = S.SurrogateParams.random_init(L.local_offsets(3), H=20, seed=10)
    ch.attach(Hp.tables_from_params(lat, p))
    of = ch.objective()
    gather = P.BestGather(ch, chain_id0=1000, world=1)
    best_of, best_chain, bits = gather.gather()
    assert best_of == of.max() and best_chain == 1000 + int(np.argmax(of))
    assert (eng.unpack_bits(bits[None, :], 144)[0] == x[int(np.argmax(of))]).all()
    of_dev = torch.zeros(n, dtype=torch.float64, device="cuda")
    B.check(ch.lib.so_objectives_dev(ch.h, C.c_void_p(of_dev.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert (of_dev.cpu().numpy() == of).all()
    pt = P.ReplicaExchange(ch, ladder=6, seed=3, rank=0, world=1)
    ms = pt.round(0, t_ref=0.05)
    assert ms >= 0 and sorted(ch.get_tscale().tolist()) == sorted(P.ladder_tscale(np.arange(n), 6).tolist())


@pytest.mark.parametrize("d,descriptor,gated,variant", [(20, "local", False, 3), (12, "local", False, 4), (12, "local", False, 0),
                                                       (12, "full", True, 4), (12, "full", True, 1), (12, "full", False, 2),
                                                       (24, "local", True, 3), (12, "full", True, 5)])
def test_type_histograms_tracked_inside_the_sa_step(eng, d, descriptor, gated, variant):
    """so_track_types: the LSC / exported-NH3 histograms every SA kernel carries through its accepted flips equal a fresh
    typing of the final structures (and the oracle's), bit for bit -- window, resident, team, flat, gated and cached-gate
    kernels; hot chains (about a third of the proposals accepted)."""
    N, n, steps = d * d, 300, 700
    if descriptor == "full":
        p = Hp.golden_surrogate(gated=gated)
    else:
        p = S.SurrogateParams.random_init(L.local_offsets(3), H=20, seed=4)
        if gated:
            rng = np.random.default_rng(1)
            feat, thr, left, right, act = [], [], [], [], []

            def node(level):
                i = len(feat)
                feat.append(-2); thr.append(0.0); left.append(-1); right.append(-1); act.append(int(rng.random() < 0.6))
                if level < 5:
                    feat[i] = int(rng.integers(0, 49)); thr[i] = 0.5
                    left[i] = node(level + 1); right[i] = node(level + 1)
                return i
            node(0)
            p = S.SurrogateParams(p.offsets, p.W1, p.b1, p.w2, p.b2, p.y_mean, p.y_scale,
                                  dict(feature=np.array(feat, np.int32), thr=np.array(thr), left=np.array(left, np.int32),
                                       right=np.array(right, np.int32), active=np.array(act, np.uint8)))
    lat = eng.Lattice(d, "NH3")
    ch = eng.Chains(lat, n if variant != 5 else 8)
    ch.randomize(seed=9, coverage=0.6)
    ch.attach(Hp.tables_from_params(lat, p))
    ch.track_types(True)
    x0 = ch.get_occupancy()
    lh, ne = ch.type_histograms()
    t0 = ST.nh3_types(x0, d)
    assert (ne == ST.nh3_export_hist(t0)).all() and (lh == np.stack([(ST.lsc_types(x0, d) == t).sum(1) for t in range(3)], 1)).all()
    T_hot = 3.0 * np.abs(ch.objective()).mean() / N if descriptor == "full" else 1.0
    st = ch.anneal(np.full(steps, T_hot), seed=5, variant=variant)
    assert st["accepted"] > 0.1 * st["flips"]
    x1 = ch.get_occupancy()
    lh, ne = ch.type_histograms()
    fresh = ch.descriptors(lsc_hist=True, nh3_exp=True)
    assert (lh == fresh["lsc_hist"]).all() and (ne == fresh["nh3_exp"]).all()
    assert (ne == ST.nh3_export_hist(ST.nh3_types(x1, d))).all()
    # explicit flips and re-assignments refresh the tracked counts
    ch.flip([0, 1], [3, 7])
    ch.set_occupancy(x0[2:3], chain0=2)
    lh, ne = ch.type_histograms()
    fresh = ch.descriptors(lsc_hist=True, nh3_exp=True)
    assert (lh == fresh["lsc_hist"]).all() and (ne == fresh["nh3_exp"]).all()
    ch.track_types(False)
    with pytest.raises(ValueError):
        ch.type_histograms()


@pytest.mark.parametrize("d,n,steps,T0", [(7, 40, 600, 2.0), (12, 300, 800, 1.5), (20, 300, 600, 1.0), (96, 64, 1200, 1.0)])
def test_compact_state_window_kernel(eng, d, n, steps, T0):
    """24-bit packed first-layer state (so_surrogate_create_compact, 64 bytes per site): the window kernel on it takes the
    decisions of the exact integer arithmetic on the coarser grid -- oracle.quantize(bits='packed24') -- bit for bit: accept
    vectors, occupancies, exact sums, the whole unpacked state; scores within the 1e-5 bar of the fp64 surrogate; hot chains
    (boundary-crossing and stale windows everywhere on the small lattices)."""
    from oracle import c_oracle
    N = d * d
    p = S.SurrogateParams.random_init(L.local_offsets(3), H=20, seed=21, y_mean=0.1, y_scale=1.7)
    q = S.quantize(p, bits="packed24")
    lat = eng.Lattice(d, "NH3")
    tab = eng.SurrogateTables(lat, p.offsets, p.W1, p.b1, p.w2, p.b2, p.y_mean, p.y_scale, compact=True)
    assert tab.e_h == q["e_h"] and tab.e_w == q["e_w"]
    rng = np.random.default_rng(d)
    x0 = (rng.random((n, N)) < 0.5).astype(np.uint8)
    prop = rng.integers(0, N, (n, steps)).astype(np.int32)
    u = rng.random((n, steps), dtype=np.float32)
    temps = SA.schedule("exp", T0, steps)
    ch = eng.Chains(lat, n)
    ch.set_occupancy(x0)
    ch.attach(tab)
    hq0, _ = S.preact_fixed(p, q, x0[3], d)
    assert (ch.preact(3)[:, :20] == hq0[:, :20]).all()                   # pack / unpack round trip of the state build
    # the tensor-core scorer packs in its epilogue; the two-pass build (int32 slab + pack kernel) must give the same state
    import os
    of_fused, pre_fused = ch.objective(), ch.preact(n - 1)
    os.environ["SO_SCORE_PACK"] = "0"
    try:
        assert (ch.score() == of_fused).all() and (ch.preact(n - 1) == pre_fused).all()
    finally:
        del os.environ["SO_SCORE_PACK"]
    assert (ch.score() == of_fused).all() and (ch.preact(n - 1) == pre_fused).all()
    for c in range(0, n, max(1, n // 8)):
        ref = S.eval_rate_fp64(p, L.descriptor_rows(x0[c], p.offsets, d))
        assert abs(ch.objective(c, 1)[0] - ref) <= 1e-5 * max(abs(ref), 1e-12)
    ch.track_types(True)
    st, acc = ch.anneal(temps, proposals=prop, uniforms=u, want_accept=True)
    assert st["kernel"] == "sa_window_kernel" and st["accepted"] > 0.05 * st["flips"]
    o = c_oracle.anneal_fixed_many(p, q, d, x0, temps, prop, u)
    assert (acc.astype(bool) == o["accept"]).all()
    x1 = ch.get_occupancy()
    assert (x1 == o["x"]).all()
    of, hi, lo, na = ch.objective(sums=True)
    assert (hi == o["hi"]).all() and (lo == o["lo"]).all() and (of == o["OF"]).all()
    hq1, _ = S.preact_fixed(p, q, x1[5], d)
    assert (ch.preact(5)[:, :20] == hq1[:, :20]).all()                   # every accepted update landed, repacked exactly
    assert (ch.score() == of).all()                                      # tracked sums == fresh recompute
    lh, ne = ch.type_histograms()
    assert (ne == ST.nh3_export_hist(ST.nh3_types(x1, d))).all()
    # the same steps in three launches (resume) and from Philox draws twice (determinism)
    ch.set_occupancy(x0)
    for a, b in ((0, 100), (100, 101), (101, steps)):
        ch.anneal(temps[a:b], step0=a, proposals=prop[:, a:b], uniforms=u[:, a:b])
    assert (ch.get_occupancy() == x1).all()
    ch.set_occupancy(x0)
    ch.anneal(temps, seed=3)
    xa = ch.get_occupancy()
    ch.set_occupancy(x0)
    ch.anneal(temps, seed=3)
    assert (ch.get_occupancy() == xa).all()
    with pytest.raises(ValueError):
        ch.anneal(temps[:10], seed=3, variant=2)                         # the generic kernels read the int32 layout
    with pytest.raises(ValueError):
        eng.SurrogateTables(lat, p.offsets, p.W1[:, :8], p.b1[:8], p.w2[:8], p.b2, compact=True)
    # near ties are re-decided from the fp64 weights: decisions equal the fp64 rule
    ch.set_occupancy(x0)
    tol = 8.0 * np.sqrt(tab.K) * np.ldexp(1.0, tab.e_h) * abs(p.y_scale) * float(np.abs(p.w2).sum())
    _, acc2 = ch.anneal(temps, proposals=prop, uniforms=u, want_accept=True, exact_guard=True, guard_tol=tol)
    o64 = c_oracle.anneal_fp64_many(p, d, x0, temps, prop, u, margin_thr=1e-9)
    ok = o64["n_below"] == 0
    assert (acc2.astype(bool)[ok] == o64["accept"][ok]).all() and ok.mean() > 0.95


@pytest.mark.parametrize("descriptor,variants", [("full", (0, 1, 2, 3, 4, 5)), ("local", (0, 2, 3, 4))])
def test_near_tie_guard_is_one_predicate_in_every_kernel(eng, descriptor, variants):
    """the exact guard must fire on the same steps in every SA kernel.  With a guard tolerance larger than any
    |delta - T ln u| EVERY decision at T > 0, u > 0 is re-taken from the fp64 weights, whatever the sign of the quantised
    delta: guard_fired counts all of them in each kernel, and the accept vectors equal the fp64 rule's."""
    from oracle import c_oracle
    d, N, n, steps = 12, 144, 48, 300
    gated = descriptor == "full"
    p = Hp.golden_surrogate(gated=True) if gated else S.SurrogateParams.random_init(L.local_offsets(3), H=20, seed=8)
    lat = eng.Lattice(d, "LSC")
    tab = Hp.tables_from_params(lat, p)
    rng = np.random.default_rng(12)
    x0 = (rng.random((n, N)) < 0.5).astype(np.uint8)
    prop = rng.integers(0, N, (n, steps)).astype(np.int32)
    u = rng.random((n, steps), dtype=np.float32)
    u[u == 0] = 0.5
    temps = SA.schedule("exp", 0.3, steps)
    o64 = c_oracle.anneal_fp64_many(p, d, x0, temps, prop, u, margin_thr=1e-9)
    assert o64["n_below"].sum() == 0
    for v in variants:
        ch = eng.Chains(lat, n if v != 5 else 4)
        m = ch.n
        ch.set_occupancy(x0[:m])
        ch.attach(tab)
        st, acc = ch.anneal(temps, proposals=prop[:m], uniforms=u[:m], want_accept=True, exact_guard=True, guard_tol=1e30,
                            variant=v)
        assert st["guard_fired"] == m * steps, (v, st["kernel"], st["guard_fired"])
        assert (acc.astype(bool) == o64["accept"][:m]).all(), (v, st["kernel"])
        assert (ch.get_occupancy() == o64["x"][:m]).all()
        ch.close()


def test_checkpoint_resume_is_bit_exact(eng, tmp_path):
    """so_chains_save / so_chains_load: a run cut at a checkpoint and resumed in a NEW handle ends with the same bits,
    sums and temperatures as the uninterrupted run (counter-based draws; the state is an exact function of the bits)."""
    d, N, n = 20, 400, 500
    p = S.SurrogateParams.random_init(L.local_offsets(3), H=20, seed=4)
    lat = eng.Lattice(d, "NH3")
    tab = Hp.tables_from_params(lat, p)
    temps = SA.schedule("exp", 0.5, 900)
    a = eng.Chains(lat, n)
    a.randomize(seed=3, coverage=0.5)
    a.attach(tab)
    a.set_tscale(np.linspace(0.5, 2.0, n))
    a.anneal(temps[:400], step0=0, seed=17, chain_id0=40)
    path = str(tmp_path / "chains.ckpt")
    a.save(path, step=400, seed=17)
    a.anneal(temps[400:], step0=400, seed=17, chain_id0=40)
    b = eng.Chains(lat, n)
    b.attach(tab)
    step, seed = b.load(path)
    assert (step, seed) == (400, 17)
    b.anneal(temps[step:], step0=step, seed=seed, chain_id0=40)
    assert (a.get_bits() == b.get_bits()).all() and (a.get_tscale() == b.get_tscale()).all()
    oa, ob = a.objective(sums=True), b.objective(sums=True)
    assert all((u == v).all() for u, v in zip(oa, ob))
    assert (a.preact(123) == b.preact(123)).all()
    # shape and surrogate mismatches are refused
    c = eng.Chains(lat, n + 1)
    with pytest.raises(ValueError):
        c.load(path)
    q = S.SurrogateParams.random_init(L.local_offsets(3), H=20, seed=5)
    e = eng.Chains(lat, n)
    e.attach(Hp.tables_from_params(lat, q))
    with pytest.raises(ValueError):
        e.load(path)
    with pytest.raises(ValueError):
        e.load(str(tmp_path / "missing.ckpt"))


# ---- BASELINE configs[1] at full size: 4,096 chains x 14,400 steps, injected proposal/uniform sequence -----------
def _config2(eng, p, n_chains):
    from oracle import c_oracle
    d, N, steps = 12, 144, 14400                                  # n_cycles = 100 (optimize_SA.py:11,25)
    rng = np.random.Generator(np.random.Philox(key=1234))
    x0 = (rng.random((n_chains, N)) < 0.5).astype(np.uint8)
    prop = rng.integers(0, N, size=(n_chains, steps)).astype(np.int32)
    u = rng.random((n_chains, steps), dtype=np.float32)
    temps = SA.schedule("exp", 0.05, steps)
    lat = eng.Lattice(d, "NH3")
    ch = eng.Chains(lat, n_chains)
    ch.set_occupancy(x0)
    ch.attach(Hp.tables_from_params(lat, p))
    stats, acc = ch.anneal(temps, proposals=prop, uniforms=u, want_accept=True)
    o = c_oracle.anneal_fixed_many(p, S.quantize(p), d, x0, temps, prop, u)
    assert (acc.astype(bool) == o["accept"]).all()                # accept/reject bit-vector identical
    x1 = ch.get_occupancy()
    assert (x1 == o["x"]).all()                                   # final occupancies identical
    of, hi, lo, na = ch.objective(sums=True)
    assert (hi == o["hi"]).all() and (lo == o["lo"]).all() and (na == o["n_act"]).all() and (of == o["OF"]).all()
    # descriptors of the final structures, bit-exact
    out = ch.descriptors(c6=True, lsc_type=True, nh3_exp=True)
    assert (out["c6"] == ST.c6_counts(x1, d)).all() and (out["lsc_type"] == ST.lsc_types(x1, d)).all()
    assert (out["nh3_exp"] == ST.nh3_export_hist(ST.nh3_types(x1, d))).all()
    # scores of a sample against the reference's fp64 arithmetic
    for c in range(0, n_chains, max(1, n_chains // 16)):
        ref = S.eval_rate_fp64(p, L.generate_all_translations(x1[c], d))
        assert abs(of[c] - ref) <= 1e-5 * max(abs(ref), 1e-12)
    assert stats["flips"] == n_chains * steps


def test_config2_full_size_open(eng):
    _config2(eng, Hp.golden_surrogate(gated=False), 4096)


def test_config2_gated(eng):
    _config2(eng, Hp.golden_surrogate(gated=True), 256)


def _config2_fp64_rule(eng, p, n_chains, y_scale_w2sum):
    """BASELINE configs[1] at full length against the REFERENCE'S fp64 rule (original weights, oracle_anneal_fp64 -- which
    tests/test_oracle_c.py pins to the reference-executed optimize() runs), exact_guard on: accept vectors and occupancies
    identical; the minimum Metropolis margin |exp(delta/T) - u| and the number of decisions closer than 1e-9 are printed.
    Every 1,440 steps: Ni-neighbour counts, LSC / NH3 types and histograms of ALL chains against
    the oracle's typing of the oracle's own occupancies."""
    from oracle import c_oracle
    d, N, steps, seg = 12, 144, 14400, 1440
    rng = np.random.Generator(np.random.Philox(key=1234))
    x0 = (rng.random((n_chains, N)) < 0.5).astype(np.uint8)
    prop = rng.integers(0, N, size=(n_chains, steps)).astype(np.int32)
    u = rng.random((n_chains, steps), dtype=np.float32)
    temps = SA.schedule("exp", 0.05, steps)
    lat = eng.Lattice(d, "NH3")
    tab = Hp.tables_from_params(lat, p)
    ch = eng.Chains(lat, n_chains)
    ch.set_occupancy(x0)
    ch.attach(tab)
    tol = 8.0 * np.sqrt(tab.K) * np.ldexp(1.0, tab.e_h) * y_scale_w2sum
    x_ref = x0.copy()
    min_margin, n_below, guard_fired = np.inf, 0, 0
    tied = np.zeros(n_chains, dtype=bool)                         # chains that met a decision closer than 1e-9
    for a in range(0, steps, seg):
        b = a + seg
        st, acc = ch.anneal(temps[a:b], step0=a, proposals=prop[:, a:b], uniforms=u[:, a:b], want_accept=True,
                            exact_guard=True, guard_tol=tol)
        guard_fired += st["guard_fired"]
        o = c_oracle.anneal_fp64_many(p, d, x_ref, temps[a:b], prop[:, a:b], u[:, a:b], margin_thr=1e-9)
        x_ref = o["x"]
        min_margin = min(min_margin, float(o["min_margin"].min()))
        n_below += int(o["n_below"].sum())
        tied |= o["n_below"] > 0
        ok = ~tied
        assert (acc.astype(bool)[ok] == o["accept"][ok]).all(), "decisions differ from the fp64 rule in steps %d..%d" % (a, b)
        x1 = ch.get_occupancy()
        assert (x1[ok] == x_ref[ok]).all()
        out = ch.descriptors(c6=True, lsc_type=True, nh3_type=True, lsc_hist=True, nh3_hist=True, nh3_exp=True)
        t_ref = ST.nh3_types(x_ref, d)
        l_ref = ST.lsc_types(x_ref, d)
        assert (out["c6"][ok] == ST.c6_counts(x_ref, d)[ok]).all() and (out["lsc_type"][ok] == l_ref[ok]).all()
        assert (out["nh3_type"][ok] == t_ref[ok]).all()
        assert (out["nh3_hist"][ok] == ST.nh3_hist(t_ref)[ok]).all() and (out["nh3_exp"][ok] == ST.nh3_export_hist(t_ref)[ok]).all()
        assert (out["lsc_hist"][ok] == np.stack([(l_ref == t).sum(1) for t in range(3)], 1)[ok]).all()
        x_ref = np.where(tied[:, None], x1, x_ref)                # a tied chain continues from the device's state
    of = ch.objective()
    assert np.abs(of[~tied] - o["OF"][~tied]).max() <= 1e-5 * np.abs(o["OF"]).max()
    print("config 2 vs the fp64 rule: %d chains x %d steps, min margin |exp(delta/T)-u| = %.3e, %d decisions below 1e-9 "
          "(%d chains excluded), exact guard fired %d times" % (n_chains, steps, min_margin, n_below, int(tied.sum()), guard_fired))
    assert tied.sum() <= max(1, n_chains // 1000)


def test_config2_fp64_rule_full_size_open(eng):
    p = Hp.golden_surrogate(gated=False)
    _config2_fp64_rule(eng, p, 4096, abs(p.y_scale) * float(np.abs(p.w2).sum()))


def test_config2_fp64_rule_gated(eng):
    p = Hp.golden_surrogate(gated=True)
    _config2_fp64_rule(eng, p, 512, abs(p.y_scale) * float(np.abs(p.w2).sum()))


# ---- BASELINE configs[3] shape (96x96, local 7x7 MLP): oracle parity on a sample + size-independent properties -----
def test_config4_lattice_96(eng):
    from oracle import c_oracle
    d, N, n_chains, steps = 96, 9216, 2048, 1500
    p = S.SurrogateParams.random_init(L.local_offsets(3), H=20, seed=0)
    q = S.quantize(p)
    lat = eng.Lattice(d, "NH3")
    tab = Hp.tables_from_params(lat, p)
    ch = eng.Chains(lat, n_chains)
    ch.randomize(seed=1234, coverage=0.5)
    x0 = ch.get_occupancy()
    ch.attach(tab)
    temps = SA.schedule("exp", 0.3, steps)
    seed, chain_id0, step0 = 99, 70000, 11
    stats, acc = ch.anneal(temps, step0=step0, seed=seed, chain_id0=chain_id0, want_accept=True)
    x1 = ch.get_occupancy()
    of, hi, lo, na = ch.objective(sums=True)
    # (1) oracle parity on a sample of chains, with the Philox draws restated by the oracle
    sample = np.arange(0, n_chains, 97)
    prop, u = philox.draws(seed, (chain_id0 + sample)[:, None], (step0 + np.arange(steps))[None, :], N)
    o = c_oracle.anneal_fixed_many(p, q, d, x0[sample], temps, prop, u)
    assert (acc[sample].astype(bool) == o["accept"]).all() and (x1[sample] == o["x"]).all()
    assert (hi[sample] == o["hi"]).all() and (lo[sample] == o["lo"]).all() and (of[sample] == o["OF"]).all()
    # (2) state invariant on ALL chains: tracked sums and first-layer state == fresh recompute from the final bits
    fresh = ch.score()
    assert (fresh == of).all()
    again = ch.objective(sums=True)
    assert (again[1] == hi).all() and (again[2] == lo).all()
    # (3) flip bookkeeping: every chain's Hamming distance to its start has the parity of its accepted-flip count
    assert ((x0 != x1).sum(1) % 2 == acc.sum(1) % 2).all()
    # (4) determinism: same seed and counters -> same bits (counter-based RNG, no atomics in the data path)
    ch.set_occupancy(x0)
    ch.anneal(temps, step0=step0, seed=seed, chain_id0=chain_id0)
    assert (ch.get_occupancy() == x1).all()
    # (5) segmentation invariance: the same steps in three launches give the same result (resume is exact)
    ch.set_occupancy(x0)
    for a, b in ((0, 400), (400, 401), (401, steps)):
        ch.anneal(temps[a:b], step0=step0 + a, seed=seed, chain_id0=chain_id0)
    assert (ch.get_occupancy() == x1).all()
    # (6) site types of the final structures bit-exact on a sample, histograms consistent on all chains
    out = ch.descriptors(chain0=0, n=n_chains, lsc_hist=True, nh3_hist=True, nh3_exp=True)
    assert (out["lsc_hist"].sum(1) == N).all() and (out["nh3_hist"].sum(1) == 5 * N).all()
    assert (out["nh3_exp"][:, 1:] == out["nh3_hist"][:, [1, 3, 5, 16, 17, 18, 19]]).all()
    t = ch.descriptors(chain0=5, n=3, nh3_type=True, lsc_type=True)
    assert (t["nh3_type"] == ST.nh3_types(x1[5:8], d)).all() and (t["lsc_type"] == ST.lsc_types(x1[5:8], d)).all()


# ---- BASELINE configs[2]: batch scoring, no annealing ---------------------------------------------------------------
def test_batch_scoring_large(eng):
    d, N, n = 12, 144, 1 << 20                                     # BASELINE configs[2]: 1M structures
    rng = np.random.default_rng(2)
    x = (rng.random((n, N)) < rng.random((n, 1))).astype(np.uint8)
    x[0], x[1] = 0, 1                                               # empty and full lattice
    p = Hp.golden_surrogate(gated=True)
    lat = eng.Lattice(d, "NH3")
    tab = Hp.tables_from_params(lat, p)
    rate, lh, ne = tab.score_batch(eng.pack_bits(x), want_lsc_hist=True, want_nh3_exp=True)
    q = S.quantize(p)
    for c in list(range(0, 64)) + list(range(64, n, n // 97)):      # oracle parity on a sample, bit-exact
        assert rate[c] == S.score_fixed(p, q, x[c], d)[0]
    lsc = ST.lsc_types(x[:2048], d)
    assert (lh[:2048] == np.stack([(lsc == t).sum(1) for t in range(3)], 1)).all()
    assert (ne[:2048] == ST.nh3_export_hist(ST.nh3_types(x[:2048], d))).all()
    # size-independent properties on ALL structures: histogram totals; translation invariance (a translated structure
    # has the same multiset of descriptor rows, hence bit-identical exact sums and the same site-type histograms)
    assert (lh.sum(1) == N).all() and (ne.sum(1) == 5 * N).all()
    xs = np.roll(x.reshape(n, d, d), shift=(5, 3), axis=(1, 2)).reshape(n, N)
    rate2, lh2, ne2 = tab.score_batch(eng.pack_bits(xs), want_lsc_hist=True, want_nh3_exp=True)
    assert (rate2 == rate).all() and (lh2 == lh).all() and (ne2 == ne).all()
    ref = S.eval_rate_fp64(p, L.generate_all_translations(x[7], d))
    assert abs(rate[7] - ref) <= 1e-5 * max(abs(ref), 1e-12)


# ---- edge cases of the C ABI: empty / ragged ranges, invalid inputs, large lattices -----------------------------------
def test_abi_edge_cases(eng):
    import ctypes as C
    from so_b200 import _lib as B
    lat = eng.Lattice(12, "NH3")
    ch = eng.Chains(lat, 8)
    lib = ch.lib
    # empty ranges are no-ops
    assert lib.so_set_occupancy(ch.h, 3, 0, None) == 0 and lib.so_get_occupancy(ch.h, 8, 0, None) == 0
    assert lib.so_descriptors(ch.h, 0, 0, None, None, None, None, None, None) == 0
    assert lib.so_flip(ch.h, 0, None, None) == 0
    # all outputs NULL: still valid
    assert lib.so_descriptors(ch.h, 0, 8, None, None, None, None, None, None) == 0
    # ranges past the end / negative
    occ = np.zeros((8, 144), dtype=np.uint8)
    for c0, n in ((1, 8), (-1, 2), (9, 0)):
        assert lib.so_get_occupancy(ch.h, c0, n, occ.ctypes.data_as(B.c_u8p)) == B.SO_EINVAL
    # calls that need a surrogate before one is attached
    of = np.zeros(8)
    assert lib.so_score(ch.h, 0, 8, of.ctypes.data_as(B.c_f64p)) == B.SO_EINVAL
    a = B.AnnealArgs()
    temps = np.ones(4)
    a.n_steps, a.temps = 4, temps.ctypes.data_as(B.c_f64p)
    assert lib.so_anneal(ch.h, C.byref(a), None) == B.SO_EINVAL and b"surrogate" in lib.so_last_error()
    # bad surrogate shapes / aliasing offsets / tree out of range
    p = S.SurrogateParams.random_init(L.local_offsets(3), H=20, seed=1)
    with pytest.raises(ValueError):
        eng.SurrogateTables(eng.Lattice(6, "LSC"), p.offsets, p.W1, p.b1, p.w2, p.b2)     # 7x7 window aliases on 6x6
    with pytest.raises(ValueError):
        eng.SurrogateTables(lat, p.offsets, np.zeros((49, 40)), np.zeros(40), np.zeros(40), 0.0)   # H > 32
    bad_tree = dict(feature=np.array([60, -2, -2]), thr=np.array([0.5, 0, 0]), left=np.array([1, -1, -1]),
                    right=np.array([2, -1, -1]), active=np.array([0, 1, 0]))
    with pytest.raises(ValueError):
        eng.SurrogateTables(lat, p.offsets, p.W1, p.b1, p.w2, p.b2, tree=bad_tree)        # feature 60 >= K
    tab = Hp.tables_from_params(lat, p)
    ch.attach(tab)
    # proposals out of range / proposals without uniforms
    prop = np.full((8, 4), 144, dtype=np.int32)
    with pytest.raises(ValueError):
        ch.anneal(temps, proposals=prop, uniforms=np.zeros((8, 4), dtype=np.float32))
    a.proposals = prop.ctypes.data_as(B.c_i32p)
    assert lib.so_anneal(ch.h, C.byref(a), None) == B.SO_EINVAL
    # zero steps: nothing happens, stats are zero
    st = ch.anneal(np.zeros(0))
    assert st["flips"] == 0 and st["accepted"] == 0
    # ragged update keeps untouched chains' state
    x = Hp.random_structures(8, 144, seed=2)
    ch.set_occupancy(x)
    of0 = ch.objective()
    ch.set_occupancy(1 - x[2:5], chain0=2)
    of1 = ch.objective()
    assert (of1[[0, 1, 5, 6, 7]] == of0[[0, 1, 5, 6, 7]]).all() and (ch.score() == of1).all()
    # exchange argument validation
    import torch
    buf = torch.zeros(8, dtype=torch.float64, device="cuda")
    ns = C.c_int64()
    for rank, world, ladder, parity in ((0, 1, 3, 0), (1, 1, 4, 0), (0, 1, 4, 2), (0, 1, 1, 0)):
        rc = lib.so_exchange(ch.h, rank, world, ladder, parity, 1, 0, 0.05, C.c_void_p(buf.data_ptr()),
                             C.c_void_p(buf.data_ptr()), None, C.byref(ns))
        assert rc == B.SO_EINVAL


def test_large_lattice_descriptors_and_flat_fallback(eng):
    """d = 160 (25,600 sites): typing against the oracle, and SA through the generic kernel (the bit-plane of such a
    lattice does not fit the pipelined kernel's shared-memory budget at 8 warps per CTA)."""
    d = 160
    N = d * d
    x = Hp.random_structures(3, N, seed=8)
    lat = eng.Lattice(d, "NH3")
    ch = eng.Chains(lat, 3)
    ch.set_occupancy(x)
    out = ch.descriptors(c6=True, lsc_type=True, nh3_type=True, nh3_exp=True)
    assert (out["c6"] == ST.c6_counts(x, d)).all() and (out["lsc_type"] == ST.lsc_types(x, d)).all()
    nh3 = ST.nh3_types(x, d)
    assert (out["nh3_type"] == nh3).all() and (out["nh3_exp"] == ST.nh3_export_hist(nh3)).all()
    p = S.SurrogateParams.random_init(L.local_offsets(3), H=20, seed=2)
    ch.attach(Hp.tables_from_params(lat, p))
    steps = 300
    rng = np.random.default_rng(1)
    prop = rng.integers(0, N, (3, steps)).astype(np.int32)
    u = rng.random((3, steps), dtype=np.float32)
    temps = SA.schedule("exp", 0.5, steps)
    stats, acc = ch.anneal(temps, proposals=prop, uniforms=u, want_accept=True)
    from oracle import c_oracle
    o = c_oracle.anneal_fixed_many(p, S.quantize(p), d, x, temps, prop, u)
    assert (acc.astype(bool) == o["accept"]).all() and (ch.get_occupancy() == o["x"]).all()
    assert (ch.objective() == o["OF"]).all()


@pytest.mark.parametrize("d", [4, 5, 7, 12, 16, 31, 32, 33, 40, 64, 96])
def test_descriptor_histograms_bit_sliced(eng, d):
    """Histogram-only requests run the bit-sliced kernel (32 sites per integer op, popc): bit-exact vs the oracle for
    lattice edges below, at and above the 32-bit word size, incl. the non-periodic type-9 rule near the cell edge."""
    N = d * d
    n = 24 if d >= 64 else 80
    x = np.vstack([Hp.random_structures(n, N, seed=300 + d), np.zeros((1, N), np.uint8), np.ones((1, N), np.uint8),
                   Hp.random_structures(8, N, seed=301 + d, coverage=0.08), Hp.random_structures(8, N, seed=302 + d, coverage=0.93)])
    lat = eng.Lattice(d, "NH3")
    ch = eng.Chains(lat, len(x))
    ch.set_occupancy(x)
    out = ch.descriptors(lsc_hist=True, nh3_hist=True, nh3_exp=True)
    lsc = ST.lsc_types(x, d)
    nh3 = ST.nh3_types(x, d)
    assert (out["lsc_hist"] == np.stack([(lsc == t).sum(1) for t in range(3)], 1)).all()
    want = ST.nh3_hist(nh3)
    assert (out["nh3_hist"] == want).all(), np.argwhere(out["nh3_hist"] != want)[:5]
    assert (out["nh3_exp"] == ST.nh3_export_hist(nh3)).all()
    os.environ["SO_DESC_BITS"] = "0"                              # the per-site kernel gives the same histograms
    try:
        ref = ch.descriptors(lsc_hist=True, nh3_hist=True, nh3_exp=True)
    finally:
        del os.environ["SO_DESC_BITS"]
    assert all((ref[k] == out[k]).all() for k in out)
    lsc_only = eng.Chains(eng.Lattice(d, "LSC"), len(x))
    lsc_only.set_occupancy(x)
    assert (lsc_only.descriptors(lsc_hist=True)["lsc_hist"] == out["lsc_hist"]).all()
