"""
ORACLE (test infrastructure, NOT product code) -- site typing of the LSC and NH3 catalysts as closed-form
local rules on the layer-3 occupancy, vectorised over a batch of structures.

Reference followed:
  OML/LSC_cat.py:176-245   types {None,1,2} + KMC_lat.site_type_inds / var_ind_kmc_sites export
  OML/NH3_cat.py:169-426   19 site types (later rules overwrite earlier ones) + the 7 exported types
The reference finds these with NetworkX VF2 sub-graph matching on the 5N-node slab graph; the closed forms
below are checked against a literal NetworkX evaluation of the same patterns in tests/test_oracle_types.py
(third-party: networkx, un-pinned upstream; 3.6.1 here).  0 encodes the reference's `None`.
"""
import numpy as np
from . import lattice as L

NH3_EXPORT = {1: 1, 3: 2, 5: 3, 16: 4, 17: 5, 18: 6, 19: 7}     # NH3_cat.py:421


def _grid(x, d):
    x = np.asarray(x)
    if x.ndim == 1:
        x = x[None, :]
    return x.reshape(x.shape[0], d, d).astype(np.int32)          # [B, i2, i1]


def _at(g, o1, o2):
    """value at site (i1+o1, i2+o2) for every (i1, i2)."""
    return np.roll(g, shift=(-o2, -o1), axis=(1, 2))


def ring_bits(g):
    return [_at(g, a, b) for (a, b) in L.RING]


def c6_counts(x, d):
    """Number of Ni among the 6 in-plane neighbours (LSC_cat.py:186-196).  [B, N] uint8."""
    g = _grid(x, d)
    return sum(ring_bits(g)).reshape(g.shape[0], -1).astype(np.uint8)


def _adjacent_vacancy_pair(r):
    out = np.zeros_like(r[0], dtype=bool)
    for k in range(6):
        out |= (r[k] == 0) & (r[(k + 1) % 6] == 0)
    return out


def lsc_types(x, d):
    """[B, N] uint8 in {0,1,2} for the layer-3 sites (all other layers are None in LSC)."""
    g = _grid(x, d)
    r = ring_bits(g)
    c6 = sum(r)
    t = np.zeros_like(g)
    t[(g == 1) & (c6 == 6)] = 1
    t[(g == 1) & (c6 == 4) & _adjacent_vacancy_pair(r)] = 2
    return t.reshape(g.shape[0], -1).astype(np.uint8)


def lsc_export(x, d):
    """(site_type_inds, var_ind_kmc_sites) of ONE structure, ascending atom index (LSC_cat.py:235-241)."""
    t = lsc_types(x, d)[0]
    sites = np.nonzero(t)[0]
    return t[sites].astype(int).tolist(), sites.astype(int).tolist()


def lsc_hist(x, d):
    t = lsc_types(x, d)
    return np.stack([(t == 1).sum(1), (t == 2).sum(1)], axis=1).astype(np.int32)


def nh3_types(x, d, h5_table=None):
    """All 19 NH3 site types over the 5N nodes: [B, 5N] uint8 (0 = None)."""
    g = _grid(x, d)
    B = g.shape[0]
    N = d * d
    r = ring_bits(g)
    c6 = sum(r)
    # layer 3, Ni: 4 (corner) unless 3 (top) or 5 (edge)
    t3 = np.where(g == 1, 4, 6).astype(np.int32)
    t3[(g == 1) & (c6 == 6)] = 3
    t3[(g == 1) & (c6 == 4) & _adjacent_vacancy_pair(r)] = 5
    is45 = ((t3 == 4) | (t3 == 5)).astype(np.int32)

    # layer 4 (vacuum nodes): depends on the 3 layer-3 sites under it
    n4 = sum(_at(g, a, b) for (a, b) in L.DN_43)
    e4 = sum(_at(is45, a, b) for (a, b) in L.DN_43)
    t4 = np.zeros_like(g)
    t4[n4 == 3] = 1
    t4[(n4 == 3) & (e4 == 1)] = 17
    t4[(n4 == 3) & (e4 >= 2)] = 16
    t4[n4 == 1] = 12
    t4[n4 == 2] = 14

    # layer 2 (Pt nodes): depends on the 3 layer-3 sites over it
    n2 = sum(_at(g, a, b) for (a, b) in L.UP_23)
    tri_t = [_at(t3, a, b) for (a, b) in L.UP_23]
    n_top = sum((q == 3).astype(np.int32) for q in tri_t)
    n_edge = sum((q == 5).astype(np.int32) for q in tri_t)
    t2 = np.zeros_like(g)
    t2[n2 == 3] = 2
    t2[(n2 == 3) & (n_top == 2) & (n_edge == 1)] = 18
    t2[(n2 == 3) & (n_top == 1) & (n_edge == 2)] = 19
    t2[n2 == 0] = 8
    t2[n2 == 2] = 15

    # layer 3, vacancies: 6, then 10 (next to a type-12 vacuum node), then 11 (type-14) overrides
    vac = (g == 0)
    has12 = np.zeros_like(vac)
    has14 = np.zeros_like(vac)
    for (a, b) in L.UP_34:
        has12 |= _at(t4, a, b) == 12
        has14 |= _at(t4, a, b) == 14
    t3[vac & has12] = 10
    t3[vac & has14] = 11

    # type 9: remaining type-6 vacancies with a type-15 layer-2 atom within 2*2.77 A of RAW xy distance
    if h5_table is None:
        h5_table = L.h5_pair_table(d)
    ptr, idx = h5_table
    t2f = t2.reshape(B, N)
    t3f = t3.reshape(B, N)
    is15 = (t2f == 15)
    near15 = np.zeros((B, N), dtype=bool)
    for i in range(N):
        js = idx[ptr[i]:ptr[i + 1]]
        if len(js):
            near15[:, i] = is15[:, js].any(1)
    t3f[(t3f == 6) & near15] = 9

    # layer 1: 7 if its 3 upper neighbours are all type 8; 13 if they are {8, 8, None}
    up_t = [_at(t2, a, b) for (a, b) in L.UP_12]
    n8 = sum((q == 8).astype(np.int32) for q in up_t)
    n0 = sum((q == 0).astype(np.int32) for q in up_t)
    t1 = np.zeros_like(g)
    t1[n8 == 3] = 7
    t1[(n8 == 2) & (n0 == 1)] = 13

    out = np.zeros((B, 5 * N), dtype=np.uint8)
    out[:, 1 * N:2 * N] = t1.reshape(B, N)
    out[:, 2 * N:3 * N] = t2f
    out[:, 3 * N:4 * N] = t3f
    out[:, 4 * N:5 * N] = t4.reshape(B, N)
    return out


def nh3_export_codes(types):
    """Map the 19-type array to the exported KMC site-type index (0 = not exported)."""
    lut = np.zeros(20, dtype=np.uint8)
    for k, v in NH3_EXPORT.items():
        lut[k] = v
    return lut[types]


def nh3_export(x, d):
    """KMC_lat.site_type_inds of ONE structure (ascending atom index; NH3_cat.py:421-426)."""
    codes = nh3_export_codes(nh3_types(x, d))[0]
    return codes[np.nonzero(codes)[0]].astype(int).tolist()


def nh3_hist(types):
    """[B, 20] int32 histogram of the 19 types (column 0 = None)."""
    B = types.shape[0]
    out = np.zeros((B, 20), dtype=np.int32)
    for b in range(B):
        out[b] = np.bincount(types[b], minlength=20)
    return out


def nh3_export_hist(types):
    """[B, 8] int32 histogram of the exported codes (column 0 = not exported)."""
    codes = nh3_export_codes(types)
    B = codes.shape[0]
    out = np.zeros((B, 8), dtype=np.int32)
    for b in range(B):
        out[b] = np.bincount(codes[b], minlength=8)
    return out
