"""CPU: closed-form site typing (oracle/site_types.py) against a literal NetworkX/VF2 evaluation of the reference's
patterns (tests/nx_literal.py) and against the committed golden vectors."""
import os
import numpy as np
import pytest
from oracle import site_types as ST
import helpers as Hp
import nx_literal as NX


@pytest.mark.parametrize("d,cov", [(4, 0.5), (5, 0.7), (6, 0.35), (6, 0.9)])
def test_closed_forms_equal_networkx(d, cov):
    x = (np.random.default_rng(int(d * 100 * cov)).random(d * d) < cov).astype(np.uint8)
    lsc = NX.lsc_types_nx(x, d)
    N = d * d
    assert (lsc[3 * N:4 * N] == ST.lsc_types(x, d)[0]).all() and lsc[:3 * N].sum() == 0 and lsc[4 * N:].sum() == 0
    assert (NX.nh3_types_nx(x, d) == ST.nh3_types(x, d)[0]).all()


def test_golden_types():
    z = np.load(os.path.join(Hp.GOLDEN, "types_golden.npz"))
    for d in (6, 8):
        assert (ST.lsc_types(z["x_d%d" % d], d) == z["lsc_d%d" % d]).all()
        assert (ST.nh3_types(z["x_d%d" % d], d) == z["nh3_d%d" % d]).all()


def test_extremes_and_exports():
    d = 8
    N = d * d
    full, empty = np.ones(N, np.uint8), np.zeros(N, np.uint8)
    assert (ST.lsc_types(full, d) == 1).all() and (ST.lsc_types(empty, d) == 0).all()
    t = ST.nh3_types(full, d)[0]
    assert (t[3 * N:4 * N] == 3).all() and (t[4 * N:] == 1).all() and (t[2 * N:3 * N] == 2).all() and (t[:2 * N] == 0).all()
    t = ST.nh3_types(empty, d)[0]
    assert (t[3 * N:4 * N] == 6).all() and (t[2 * N:3 * N] == 8).all() and (t[N:2 * N] == 7).all() and (t[4 * N:] == 0).all()
    assert ST.nh3_export(full, d) == [2] * N + [1] * N                 # ascending atom index: layer 3 then layer 4
    assert ST.lsc_export(full, d) == ([1] * N, list(range(N)))
    x = full.copy()
    x[[9, 10]] = 0                                                     # two adjacent vacancies -> four edge sites
    types_, sites = ST.lsc_export(x, d)
    assert types_.count(2) == 2 and ST.c6_counts(x, d)[0, 9] == 5
