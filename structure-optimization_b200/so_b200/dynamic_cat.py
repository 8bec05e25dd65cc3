"""
dynamic_cat -- drop-in for the reference's catalyst base class (OML/dynamic_cat.py:9-437) on the B200 engine.

Same constructor, attributes and method names; the occupancy vector stays a plain Python list on the host (callers
index, iterate and even re-assign it), and is mirrored into a bit-packed device chain whenever a GPU evaluation
needs it.  ASE is not required: the slab geometry comes from the engine's closed-form lattice tables
(so_lattice_create), exposed through a minimal `SlabTemplate` with the ASE accessors the reference's callers use.
"""
import random
import numpy as np
from . import engine


class SlabTemplate(object):
    """Stand-in for the ASE Atoms template (OML/dynamic_cat.py:49): positions, cell, symbols, len()."""

    def __init__(self, positions, cell2d, symbols):
        self._pos = np.asarray(positions, dtype=np.float64)
        self._cell = np.zeros((3, 3))
        self._cell[:2, :2] = cell2d
        self._cell[2, 2] = self._pos[:, 2].max() + 15.0 if len(self._pos) else 30.0     # vacuum on both sides
        self._symbols = list(symbols)

    def __len__(self):
        return len(self._symbols)

    def get_positions(self):
        return self._pos.copy()

    def get_cell(self):
        return self._cell.copy()

    def get_chemical_symbols(self):
        return list(self._symbols)

    def get_scaled_positions(self, wrap=True):
        """Fractional coordinates; like ASE, the two periodic (in-plane) axes are wrapped into [0, 1)."""
        inv = np.linalg.inv(self._cell[:2, :2])
        out = np.zeros_like(self._pos)
        out[:, :2] = self._pos[:, :2] @ inv
        out[:, 2] = self._pos[:, 2] / self._cell[2, 2]
        if wrap:
            out[:, :2] %= 1.0
            out[:, :2] %= 1.0
        return out


class dynamic_cat(object):
    """Template for defected catalyst facets to be optimized (reference: OML/dynamic_cat.py:15-59)."""

    _kind = "LSC"

    def __init__(self, met_name='Pt', facet='111', dim1=12, dim2=12, fixed_layers=3, variable_layers=1, device=0):
        super(dynamic_cat, self).__init__()
        self.dim1 = dim1
        self.dim2 = dim2
        self.atoms_per_layer = self.dim1 * self.dim2
        self.atoms_template = None
        self.atoms_defected = None
        self.variable_atoms = None
        self.variable_occs = None
        self.surface_area = None
        self.defected_graph = None
        self.atom_last_moved = None
        if not (facet == '111' or facet == 111):
            if facet == '100' or facet == 100:
                raise ValueError('(100) facets are not supported by the B200 path (the hot path is fcc(111)).')
            raise ValueError(str(facet) + ' is not a valid facet.')          # OML/dynamic_cat.py:53
        if dim1 != dim2:
            raise ValueError('the B200 path needs a square cell (dim1 == dim2), as LSC_cat / NH3_cat build it')
        total_layers = fixed_layers + variable_layers
        self._device = int(device)
        self._lattice = engine.Lattice(dim1, self._kind, device=device)
        self._chain = engine.Chains(self._lattice, 1)
        self._mirrored = None
        n = self.atoms_per_layer
        pos, cell = self._lattice.positions()
        if total_layers == 5:
            symbols = [met_name] * (5 * n)
            self.atoms_template = SlabTemplate(pos, cell, symbols)
        self._cell2d = cell
        self.variable_atoms = range(fixed_layers * n, total_layers * n)      # OML/dynamic_cat.py:57
        self.variable_occs = [1 for i in self.variable_atoms]
        self.surface_area = float(abs(cell[0, 0] * cell[1, 1] - cell[0, 1] * cell[1, 0]))

    # ---- device mirror ------------------------------------------------------------------------------------
    def _occs_array(self, x=None):
        x = self.variable_occs if x is None else x
        arr = np.asarray(x)
        if arr.shape != (self.atoms_per_layer,):
            raise ValueError('occupancy vector must have %d entries' % self.atoms_per_layer)
        if not np.isin(arr, (0, 1)).all():
            raise NameError('Invalid occupancy.')                            # OML/dynamic_cat.py:185
        return arr.astype(np.uint8)

    def _push(self):
        """Mirror the host occupancy list into device chain 0 (no-op when unchanged)."""
        arr = self._occs_array()
        if self._mirrored is None or not np.array_equal(arr, self._mirrored):
            self._chain.set_occupancy(arr[None, :])
            self._mirrored = arr
        return self._chain

    # ---- occupancies (OML/dynamic_cat.py:62-84, 152-185, 223-230) --------------------------------------------
    def randomize(self, coverage=None, build_structure=False):
        self.variable_occs = [0 for i in range(len(self.variable_atoms))]
        if coverage is None:
            coverage = random.random()
        n_occupancies = int(round(coverage * len(self.variable_atoms)))
        occupied_sites = random.sample(range(len(self.variable_occs)), n_occupancies)
        for i in occupied_sites:
            self.variable_occs[i] = 1
        if build_structure:
            self.occs_to_atoms()
            self.occs_to_graph()
        return self.variable_occs

    def assign_occs(self, x):
        self.variable_occs = x
        self.occs_to_graph()
        self.occs_to_atoms()

    def occs_to_atoms(self):
        """The reference deep-copies the ASE template and deletes missing atoms; here: the list of present atoms."""
        if self.atoms_template is None:
            self.atoms_defected = None
            return
        missing = np.zeros(len(self.atoms_template), dtype=bool)
        occ = self._occs_array()
        missing[np.asarray(self.variable_atoms)] = occ == 0
        pos = self.atoms_template.get_positions()[~missing]
        sym = [s for s, m in zip(self.atoms_template.get_chemical_symbols(), missing) if not m]
        self.atoms_defected = SlabTemplate(pos, self._cell2d, sym)

    def occs_to_graph(self):
        pass                                                                 # implemented in subclasses

    def graph_to_occs(self):
        pass

    def graph_to_atoms(self):
        self.graph_to_occs()
        self.occs_to_atoms()

    def flip_atom(self, ind):
        """`ind` is an ATOM index (variable_atoms), as upstream."""
        ind = self.variable_atoms.index(ind)
        if self.variable_occs[ind] == 1:
            self.variable_occs[ind] = 0
        elif self.variable_occs[ind] == 0:
            self.variable_occs[ind] = 1
        else:
            raise NameError('Invalid occupancy.')

    def rand_move(self):
        self.atom_last_moved = random.choice(self.variable_atoms)
        self.flip_atom(self.atom_last_moved)

    def revert_last(self):
        self.flip_atom(self.atom_last_moved)

    # ---- index maps and symmetry descriptors (OML/dynamic_cat.py:233-363) ---------------------------------------
    def sym_inds_to_var_ind(self, sym_ind1, sym_ind2):
        sym_ind1 = sym_ind1 % self.dim1
        sym_ind2 = sym_ind2 % self.dim2
        return sym_ind2 * self.dim1 + sym_ind1

    def var_ind_to_sym_inds(self, var_ind):
        return var_ind % self.dim1, var_ind // self.dim1                     # py2 integer division upstream

    def _site_grid(self):
        v = np.arange(len(self.variable_atoms))
        return v % self.dim1, v // self.dim1

    def translate(self, shift1, shift2, old_vec=None):
        if old_vec is None:
            old_vec = self.variable_occs
        d1, d2 = self._site_grid()
        src = ((d2 + shift2) % self.dim2) * self.dim1 + (d1 + shift1) % self.dim1
        return np.asarray(old_vec, dtype=np.float64)[src]

    def generate_all_translations(self, local=False):
        x = np.asarray(self.variable_occs, dtype=np.float64)
        d1, d2 = self._site_grid()
        src = ((d2[None, :] + d2[:, None]) % self.dim2) * self.dim1 + (d1[None, :] + d1[:, None]) % self.dim1
        return x[src]

    def rotate(self, i, old_occs=None):
        if old_occs is None:
            old_occs = self.variable_occs
        i = i % 3
        if i == 0:
            return old_occs
        d1, d2 = self._site_grid()
        if i == 1:
            n1, n2 = -d1 - d2, d1
        else:
            n1, n2 = d2, -d1 - d2
        src = (n2 % self.dim2) * self.dim1 + n1 % self.dim1
        return np.asarray(old_occs, dtype=np.float64)[src]

    def generate_all_translations_and_rotations(self, old_vec=None):
        if old_vec is None:
            old_vec = self.variable_occs
        x = np.asarray(old_vec, dtype=np.float64)
        n = len(self.variable_atoms)
        d1, d2 = self._site_grid()
        trans = x[((d2[None, :] + d2[:, None]) % self.dim2) * self.dim1 + (d1[None, :] + d1[:, None]) % self.dim1]
        out = np.empty((3 * n, n))
        out[0::3] = trans
        out[1::3] = trans[:, (d1 % self.dim2) * self.dim1 + (-d1 - d2) % self.dim1]
        out[2::3] = trans[:, ((-d1 - d2) % self.dim2) * self.dim1 + d2 % self.dim1]
        return out

    def get_local_inds(self, shift1=0, shift2=0):
        return [self.sym_inds_to_var_ind(d1 + shift1, d2 + shift2) for d2 in range(-3, 4) for d1 in range(-3, 4)]

    def show(self, fname='structure_1', fmat='png'):
        """OML/dynamic_cat.py:366-401 renders the slab with ASE (png / xsd / povray): visualisation is out of scope."""
        raise NotImplementedError('show() needs ASE rendering, which is outside the B200 hot path')

    # ---- GA operators / distance metric (OML/dynamic_cat.py:188-220,404-437), host side: O(N) per call
    def geo_crossover(self, x1, x2, pt1=1, pt2=1):
        x_bounds = [random.random() for i in range(pt1)]
        y_bounds = [random.random() for i in range(pt2)]
        frac = self.atoms_template.get_scaled_positions()
        for i in range(len(x1)):
            score = sum(frac[self.variable_atoms[i], 0] > b for b in x_bounds)
            score += sum(frac[self.variable_atoms[i], 1] > b for b in y_bounds)
            if score % 2 == 0:
                x1[i], x2[i] = x2[i], x1[i]
        return x1, x2

    def calc_weights(self):
        w = np.zeros(len(self.variable_occs))
        d = self.dim1
        images = [[0, 0], [0, d], [d, 0], [d, d], [d, -d]]
        for i in range(len(self.variable_occs)):
            f1, f2 = self.var_ind_to_sym_inds(i)
            moves = []
            for im in images:
                a, b = f1 - im[0], f2 - im[1]
                moves.append(min(abs(a) + abs(b), abs(a + b) + abs(a), abs(a + b) + abs(b)))
            w[i] = np.exp(-2 * min(moves))
        return w


def dyn_cat_dist(occs1, occs2, w):
    return np.dot(np.abs(occs1 - occs2), w)
