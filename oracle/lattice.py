"""
ORACLE (test infrastructure, NOT product code) -- slab geometry, neighbour topology and the
index/translation helpers of the reference catalyst base class, restated in NumPy.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this
package.  The product path (structure-optimization_b200/) never does.

Reference followed (paths relative to the upstream repo):
  OML/dynamic_cat.py:15-59    slab = fcc111(size=(d,d,5), a=3.9239, vacuum=15), PBC in-plane only
  OML/dynamic_cat.py:233-363  translate / rotate / local window / index maps
  OML/LSC_cat.py:56-79        graph = ASE NeighborList with radius (2.77+0.2)/2 per atom
Third-party arithmetic not in the reference tree: ASE `fcc111` + `NeighborList` (un-pinned, absent here).
Their output is pure geometry; it is pinned by the reference's only fixture, zacros_input/lattice_input.dat
(cell vectors + 14 fractional site coordinates + 2 neighbour pairs), which tests/test_oracle_lattice.py
checks against tests/golden/lattice_input_fixture.json.
"""
import numpy as np

A_LAT = 3.9239                      # dynamic_cat.py:49
D_NN = A_LAT / np.sqrt(2.0)         # nearest-neighbour distance
DZ = A_LAT / np.sqrt(3.0)           # interlayer spacing
VACUUM = 15.0                       # dynamic_cat.py:49
PAIR_CUTOFF = 2.77 + 0.2            # LSC_cat.py:72  (two radii of (2.77+0.2)/2)
N_LAYERS = 5                        # fixed_layers=4 + variable_layers=1 (LSC_cat.py:25)
VAR_LAYER = 3                       # LSC_cat.py:49 : variable atoms are 3N..4N-1

# In-plane position of layer l in lattice coordinates is (i1 + ox, i2 + oy); ABC stacking counted from the
# top layer.  Pinned by the fixture: the layer-3 sites only land on integers with the (-1/3, +2/3) shift.
LAYER_SHIFT = {4: (0.0, 0.0), 3: (-1.0 / 3, 2.0 / 3), 2: (1.0 / 3, 1.0 / 3), 1: (0.0, 0.0), 0: (-1.0 / 3, 2.0 / 3)}

# The six in-plane neighbours in ring order (consecutive entries are mutually adjacent).
RING = ((1, 0), (0, 1), (-1, 1), (-1, 0), (0, -1), (1, -1))
# Inter-layer neighbour offsets (delta i1, delta i2), "from layer -> to layer"
UP_34 = ((0, 0), (-1, 1), (0, 1))       # layer 3 site -> the 3 layer-4 atoms above it
DN_43 = ((0, -1), (1, -1), (0, 0))      # layer 4 atom -> the 3 layer-3 sites under it  (tri4)
DN_32 = ((-1, 0), (0, 0), (-1, 1))      # layer 3 site -> the 3 layer-2 atoms under it
UP_23 = ((1, -1), (0, 0), (1, 0))       # layer 2 atom -> the 3 layer-3 sites over it   (tri2)
UP_12 = ((-1, 0), (0, -1), (0, 0))      # layer 1 atom -> the 3 layer-2 atoms over it
DN_21 = ((1, 0), (0, 1), (0, 0))        # layer 2 atom -> the 3 layer-1 atoms under it
# layers 0/1 repeat the 3/4 registry (ABC): layer 1 sits like layer 4, layer 0 like layer 3
UP_01 = UP_34
DN_10 = DN_43


def cell_vectors(d):
    """2x2 in-plane cell (rows), as ASE builds it for the non-orthogonal fcc111 slab."""
    return np.array([[d * D_NN, 0.0], [d * D_NN * 0.5, d * D_NN * np.sqrt(0.75)]])


def slab_positions(d):
    """Raw (un-wrapped) cartesian positions of the 5*d*d template atoms, index = layer*N + i2*d + i1."""
    N = d * d
    pos = np.zeros((N_LAYERS * N, 3))
    i1 = np.tile(np.arange(d), d).astype(float)
    i2 = np.repeat(np.arange(d), d).astype(float)
    for layer in range(N_LAYERS):
        ox, oy = LAYER_SHIFT[layer]
        u, v = i1 + ox, i2 + oy
        pos[layer * N:(layer + 1) * N, 0] = D_NN * (u + 0.5 * v)
        pos[layer * N:(layer + 1) * N, 1] = D_NN * np.sqrt(0.75) * v
        pos[layer * N:(layer + 1) * N, 2] = VACUUM + layer * DZ
    return pos


def neighbour_sets_bruteforce(d):
    """Neighbour sets from distances with the in-plane minimum image (what ASE NeighborList computes,
    LSC_cat.py:71-79).  O(n^2) -- for small d only."""
    pos = slab_positions(d)
    cell = cell_vectors(d)
    n = len(pos)
    inv = np.linalg.inv(cell)
    nbrs = [set() for _ in range(n)]
    for i in range(n):
        dxy = pos[:, :2] - pos[i, :2]
        frac = dxy @ inv
        frac -= np.round(frac)
        dxy = frac @ cell
        dist = np.sqrt((dxy ** 2).sum(1) + (pos[:, 2] - pos[i, 2]) ** 2)
        for j in np.nonzero(dist < PAIR_CUTOFF)[0]:
            if j != i:
                nbrs[i].add(int(j))
    return nbrs


def _node(layer, i1, i2, d):
    return layer * d * d + (i2 % d) * d + (i1 % d)


def neighbour_csr(d):
    """Closed-form neighbour CSR (rowptr[5N+1], col[54N]); column order per node: ring, then up, then down."""
    N = d * d
    rowptr = [0]
    col = []
    up = {0: UP_01, 1: UP_12, 2: UP_23, 3: UP_34}
    dn = {1: DN_10, 2: DN_21, 3: DN_32, 4: DN_43}
    for layer in range(N_LAYERS):
        for i2 in range(d):
            for i1 in range(d):
                for (a, b) in RING:
                    col.append(_node(layer, i1 + a, i2 + b, d))
                if layer in up:
                    for (a, b) in up[layer]:
                        col.append(_node(layer + 1, i1 + a, i2 + b, d))
                if layer in dn:
                    for (a, b) in dn[layer]:
                        col.append(_node(layer - 1, i1 + a, i2 + b, d))
                rowptr.append(len(col))
    return np.asarray(rowptr, dtype=np.int32), np.asarray(col, dtype=np.int32)


# ---- index maps (dynamic_cat.py:344-363; py2 integer division made explicit) ---------------------------

def sym_inds_to_var_ind(i1, i2, d):
    return (i2 % d) * d + (i1 % d)


def var_ind_to_sym_inds(v, d):
    return v % d, v // d


def translate(x, s1, s2, d):
    """dynamic_cat.py:248-267: new[v] = old[idx(v1+s1, v2+s2)]."""
    x2 = np.asarray(x).reshape(d, d)                       # [i2, i1]
    return np.roll(x2, shift=(-s2, -s1), axis=(0, 1)).reshape(-1).astype(float)


def generate_all_translations(x, d):
    """dynamic_cat.py:233-245: row r = translate(r1, r2); returns float64 [N, N]."""
    N = d * d
    x = np.asarray(x)
    v = np.arange(N)
    v1, v2 = v % d, v // d
    out = np.empty((N, N))
    for r in range(N):
        r1, r2 = r % d, r // d
        out[r] = x[((v2 + r2) % d) * d + (v1 + r1) % d]
    return out


def rotate(x, k, d):
    """dynamic_cat.py:288-322."""
    k = k % 3
    x = np.asarray(x, dtype=float)
    if k == 0:
        return x
    v = np.arange(d * d)
    d1, d2 = v % d, v // d
    if k == 1:
        n1, n2 = -d1 - d2, d1
    else:
        n1, n2 = d2, -d1 - d2
    return x[(n2 % d) * d + (n1 % d)]


def generate_all_translations_and_rotations(x, d):
    """dynamic_cat.py:270-285: rows 3r+k."""
    N = d * d
    T = generate_all_translations(x, d)
    out = np.empty((3 * N, N))
    for k in range(3):
        v = np.arange(N)
        d1, d2 = v % d, v // d
        if k == 0:
            src = v
        elif k == 1:
            src = ((d1) % d) * d + ((-d1 - d2) % d)
        else:
            src = ((-d1 - d2) % d) * d + (d2 % d)
        out[k::3] = T[:, src]
    return out


def get_local_inds(s1, s2, d):
    """dynamic_cat.py:325-341: 7x7 window, d2 outer, d1 inner."""
    return [sym_inds_to_var_ind(a + s1, b + s2, d) for b in range(-3, 4) for a in range(-3, 4)]


def local_offsets(radius=3):
    """Descriptor column k -> lattice offset (o1, o2), in get_local_inds order."""
    return np.array([(a, b) for b in range(-radius, radius + 1) for a in range(-radius, radius + 1)], dtype=np.int32)


def full_offsets(d):
    """Descriptor column c of a translation row is the site at offset var_ind_to_sym_inds(c)."""
    v = np.arange(d * d)
    return np.stack([v % d, v // d], axis=1).astype(np.int32)


def descriptor_rows(x, offsets, d):
    """Generic descriptor matrix [N, K]: row r, column k = x[r + offsets[k]] (float64 like the reference)."""
    x = np.asarray(x)
    N = d * d
    r = np.arange(N)
    r1, r2 = r % d, r // d
    idx = ((r2[:, None] + offsets[None, :, 1]) % d) * d + (r1[:, None] + offsets[None, :, 0]) % d
    return x[idx].astype(float)


def h5_pair_table(d):
    """For NH3 type 9 (NH3_cat.py:367-381): layer-2 atoms within 2*2.77 A of each layer-3 site, measured on
    RAW (non-periodic) xy positions.  Returns CSR (ptr[N+1], idx[]) of layer-2 in-layer indices."""
    pos = slab_positions(d)
    N = d * d
    p3 = pos[3 * N:4 * N, :2]
    p2 = pos[2 * N:3 * N, :2]
    ptr = [0]
    idx = []
    for i in range(N):
        dist = np.sqrt(((p2 - p3[i]) ** 2).sum(1))
        near = np.nonzero(dist < 2.77 * 2)[0]
        idx.extend(int(j) for j in near)
        ptr.append(len(idx))
    return np.asarray(ptr, dtype=np.int32), np.asarray(idx, dtype=np.int32)
