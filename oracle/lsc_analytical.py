"""
ORACLE (test infrastructure, NOT product code) -- analytical steady state of the LSC toy model and the
synthetic-structure recipe; used to LABEL synthetic training data (BASELINE config 1) and as the oracle of
the "next" row SURVEY 8f-1.

Reference followed:
  OML/LSC_cat.py:248-316                    eval_struc_rate_anal / compute_site_rates_lsc
  drivers/generate_init_structures.py:26-116 random 19-site-island structures + 16 mutations
Third-party not in the tree: zacros_wrapper.Lattice.Build_neighbor_list(cut=2.87) (absent, un-pinned): taken
as "typed sites that are in-plane nearest neighbours under the periodic cell, each unordered pair once".
"""
import numpy as np
from . import lattice as L
from . import site_types as ST


def kmc_neighbor_pairs(sites, d):
    """Unordered pairs (i, j), i < j, of positions in `sites` that are in-plane nearest neighbours."""
    where = {int(v): i for i, v in enumerate(sites)}
    pairs = []
    for i, v in enumerate(sites):
        v1, v2 = v % d, v // d
        for (a, b) in L.RING:
            w = ((v2 + b) % d) * d + (v1 + a) % d
            j = where.get(int(w))
            if j is not None and i < j:
                pairs.append((i, j))
    return pairs


def compute_site_rates_lsc(x, d):
    """Site rates [N] of one structure (LSC_cat.py:260-316; all rate constants 1)."""
    N = d * d
    types, sites = ST.lsc_export(x, d)
    site_rates = np.zeros(N)
    n = len(sites)
    if n == 0:
        return site_rates
    A = np.zeros((n, n))
    b = np.zeros(n)
    A_rate = np.zeros((n, n))
    for i, t in enumerate(types):
        if t == 1:
            A[i, i] += -2.0
            b[i] += -1.0
        elif t == 2:
            A[i, i] += -1.0
            A_rate[i, i] = 1.0
    for (i, j) in kmc_neighbor_pairs(sites, d):
        A[i, i] -= 1.0; A[i, j] += 1.0
        A[j, j] -= 1.0; A[j, i] += 1.0
    theta = np.linalg.solve(A, b)
    rates = A_rate.dot(theta)
    for i, v in enumerate(sites):
        site_rates[v] = rates[i]
    return site_rates


def eval_struc_rate_anal(x, d):
    return float(np.sum(compute_site_rates_lsc(x, d)))


ISLAND = [(-2, 2), (-1, 2), (0, 2), (-2, 1), (-1, 1), (0, 1), (1, 1), (-2, 0), (-1, 0), (0, 0), (1, 0), (2, 0),
          (-1, -1), (0, -1), (1, -1), (2, -1), (0, -2), (1, -2), (2, -2)]


def make_random_structure(i, n_strucs, d):
    """generate_init_structures.py:40-116: islands around n_tops random sites, then 16 random mutations."""
    N = d * d
    np.random.seed(i)
    x = np.zeros(N, dtype=np.uint8)
    n_tops = int(3 + 15 * float(i) / (n_strucs + 1))
    tops = np.random.choice(range(N), size=n_tops, replace=False)
    for t in tops:
        t1, t2 = t % d, t // d
        for (a, b) in ISLAND:
            x[((t2 + b) % d) * d + (t1 + a) % d] = 1
    for site in np.random.choice(range(N), size=16, replace=False):
        x[site] ^= 1
    return x
