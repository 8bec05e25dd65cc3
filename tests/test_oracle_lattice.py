"""CPU: the oracle's geometry/topology against the reference's only fixture (zacros_input/lattice_input.dat, committed
as tests/golden/lattice_input_fixture.json) and against brute-force distances."""
import json
import os
import numpy as np
import pytest
from oracle import lattice as L
import helpers as Hp


def test_fixture_cell_and_layer3_registry():
    fx = json.load(open(os.path.join(Hp.GOLDEN, "lattice_input_fixture.json")))
    d = 12
    np.testing.assert_allclose(L.cell_vectors(d), fx["cell"], atol=6e-4)          # printed with 3 decimals
    pos = L.slab_positions(d)[3 * d * d:4 * d * d, :2]
    frac = pos @ np.linalg.inv(L.cell_vectors(d))
    # every listed KMC site is a layer-3 atom of the slab (5 printed decimals)
    sites = []
    for xy in fx["site_coordinates"]:
        dist = np.abs(frac - np.asarray(xy)).max(1)
        assert dist.min() < 2e-5
        sites.append(int(np.argmin(dist)))
    assert [(v % d, v // d) for v in sites] == [(5, 1), (11, 1), (0, 2), (4, 2), (1, 7), (8, 7), (2, 8), (8, 8), (3, 9),
                                                (6, 9), (1, 10), (2, 11), (7, 11), (10, 11)]
    # the two listed neighbour pairs are in-plane nearest neighbours of the closed-form ring
    for a, b in fx["neighbor_pairs"]:
        va, vb = sites[a - 1], sites[b - 1]
        offs = {((vb % d - va % d + d // 2) % d - d // 2, (vb // d - va // d + d // 2) % d - d // 2)}
        assert offs <= set(L.RING)


@pytest.mark.parametrize("d", [4, 5, 6, 9])
def test_closed_form_neighbours_equal_distance_search(d):
    brute = L.neighbour_sets_bruteforce(d)
    rowptr, col = L.neighbour_csr(d)
    N = d * d
    for i in range(5 * N):
        assert set(col[rowptr[i]:rowptr[i + 1]].tolist()) == brute[i]
        assert rowptr[i + 1] - rowptr[i] == (9 if i // N in (0, 4) else 12)
    assert rowptr[-1] == 54 * N


def test_translations_and_rotations_literal():
    """Vectorised helpers against the literal per-site loops of OML/dynamic_cat.py:233-341."""
    d = 6
    N = d * d
    x = (np.random.default_rng(0).random(N) < 0.5).astype(float)
    idx = lambda a, b: (b % d) * d + (a % d)
    for s1, s2 in ((0, 0), (2, 5), (-1, 3)):
        lit = np.array([x[idx(v % d + s1, v // d + s2)] for v in range(N)])
        np.testing.assert_array_equal(L.translate(x, s1, s2, d), lit)
    T = L.generate_all_translations(x, d)
    assert (T[:, 0] == x).all()                                                    # T[r, 0] = x[r]
    R = L.generate_all_translations_and_rotations(x, d)
    np.testing.assert_array_equal(R[::3], T)
    for r in (0, 7, 35):
        t = T[r]
        np.testing.assert_array_equal(R[3 * r + 1], [t[idx(-(v % d) - v // d, v % d)] for v in range(N)])
        np.testing.assert_array_equal(R[3 * r + 2], [t[idx(v // d, -(v % d) - v // d)] for v in range(N)])
    assert L.get_local_inds(0, 0, d)[24] == 0 and len(L.get_local_inds(2, 3, d)) == 49
    np.testing.assert_array_equal(L.descriptor_rows(x, L.full_offsets(d), d), T)


def test_incremental_descriptor_update_identity():
    """OML/optimize_SA.py:88-92: flipping site f changes exactly column idx(f - r) of row r."""
    d = 5
    x = (np.random.default_rng(1).random(d * d) < 0.5).astype(int)
    T = L.generate_all_translations(x, d)
    f = 13
    y = x.copy()
    y[f] ^= 1
    T2 = T.copy()
    for r in range(d * d):
        T2[r, ((f // d - r // d) % d) * d + (f % d - r % d) % d] = y[f]
    np.testing.assert_array_equal(T2, L.generate_all_translations(y, d))
