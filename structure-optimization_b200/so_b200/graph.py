"""Light-weight stand-ins for the two third-party objects the reference's catalyst classes expose:
  * `defected_graph` (a networkx.Graph upstream, OML/LSC_cat.py:58-79): here a view over the engine's CSR
    neighbour list + element / site-type arrays, with the accessors callers use (`.node[i]['element']`, ...);
  * `KMC_lat` (zacros_wrapper.Lattice upstream, OML/LSC_cat.py:228-245): site types, names, coordinates and the
    neighbour list, plus a writer for Zacros' lattice_input.dat.
"""
import numpy as np


class _NodeAttr(object):
    def __init__(self, graph, i):
        self._g, self._i = graph, i

    def __getitem__(self, key):
        if key == 'element':
            return self._g.element(self._i)
        if key == 'site_type':
            t = int(self._g.site_type[self._i])
            return t if t else None
        raise KeyError(key)

    def __setitem__(self, key, value):
        if key == 'site_type':
            self._g.site_type[self._i] = 0 if value is None else int(value)
        elif key == 'element':
            self._g.set_element(self._i, value)
        else:
            raise KeyError(key)


class _NodeView(object):
    def __init__(self, graph):
        self._g = graph

    def __getitem__(self, i):
        if not 0 <= i < self._g.n_nodes:
            raise KeyError(i)
        return _NodeAttr(self._g, i)

    def __len__(self):
        return self._g.n_nodes

    def __iter__(self):
        return iter(range(self._g.n_nodes))


class GraphView(object):
    def __init__(self, cat, rowptr, col):
        self._cat = cat
        self.rowptr, self.col = rowptr, col
        self.n_nodes = len(rowptr) - 1
        self.site_type = np.zeros(self.n_nodes, dtype=np.uint8)
        self.node = _NodeView(self)          # networkx 1.x spelling used by the reference
        self.nodes = self.node               # networkx >= 2 spelling

    def element(self, i):
        n = self._cat.atoms_per_layer
        layer = i // n
        if layer == 4:
            return 'vacuum'
        if layer == 3:
            return 'Ni' if self._cat.variable_occs[i - 3 * n] else 'vacancy'
        return 'Pt'

    def set_element(self, i, value):
        n = self._cat.atoms_per_layer
        if i // n != 3 or value not in ('Ni', 'vacancy'):
            raise NameError('Invalid element type in graph for variable atom.')
        self._cat.variable_occs[i - 3 * n] = 1 if value == 'Ni' else 0

    def neighbors(self, i):
        return iter(self.col[self.rowptr[i]:self.rowptr[i + 1]].tolist())

    def number_of_nodes(self):
        return self.n_nodes

    def number_of_edges(self):
        return len(self.col) // 2

    def edges(self):
        for i in range(self.n_nodes):
            for j in self.col[self.rowptr[i]:self.rowptr[i + 1]]:
                if i < j:
                    yield (i, int(j))


class KMCLattice(object):
    """Subset of zacros_wrapper.Lattice used on this path."""

    def __init__(self):
        self.text_only = False
        self.lattice_matrix = None
        self.site_type_names = []
        self.site_type_inds = []
        self.cart_coords = None
        self.frac_coords = None
        self.neighbor_list = []
        self.cell_list = []          # per pair: 'self' | 'north' | 'northeast' | 'east' | 'southeast'

    def set_cart_coords(self, cart_coords):
        self.cart_coords = np.asarray(cart_coords, dtype=np.float64).reshape(-1, 2)
        self.frac_coords = self.cart_coords @ np.linalg.inv(np.asarray(self.lattice_matrix))

    # Zacros names the periodic image the second site of a pair sits in (lattice_input.dat "neighboring_structure"):
    # image shift in units of the two cell vectors (a, b) -> keyword
    CELLS = ((0, 0, 'self'), (0, 1, 'north'), (1, 1, 'northeast'), (1, 0, 'east'), (1, -1, 'southeast'))

    def Build_neighbor_list(self, cut=3.0):
        """zacros_wrapper.Lattice.Build_neighbor_list: pairs [s1, s2] of sites closer than `cut`, with the cell the
        second site is taken from in `cell_list`: 'self' (s1 < s2), or s2 displaced by +b ('north'), +a+b
        ('northeast'), +a ('east'), +a-b ('southeast').  Every periodic pair appears once; a pair that crosses the
        cell edge the other way is listed with its sites exchanged."""
        self.neighbor_list, self.cell_list = [], []
        n = len(self.site_type_inds)
        if n == 0:
            return
        a, b = np.asarray(self.lattice_matrix, dtype=np.float64)
        c = self.cart_coords
        found = []
        for na, nb, name in self.CELLS:
            shift = na * a + nb * b
            dist = np.sqrt(((c[None, :, :] + shift - c[:, None, :]) ** 2).sum(2))       # [s1, s2]
            s1, s2 = np.nonzero(dist < cut)
            for i, j in zip(s1.tolist(), s2.tolist()):
                if name != 'self' or i < j:
                    found.append((i, j, name))
        order = {name: k for k, (_, _, name) in enumerate(self.CELLS)}
        found.sort(key=lambda t: (t[0], t[1], order[t[2]]))
        self.neighbor_list = [[i, j] for i, j, _ in found]
        self.cell_list = [name for _, _, name in found]

    def Write_lattice_input(self, fldr):
        """lattice_input.dat in the layout of zacros_input/lattice_input.dat:1-37."""
        import os
        n = len(self.site_type_inds)
        with open(os.path.join(fldr, 'lattice_input.dat'), 'w') as f:
            f.write('# Lattice specification file: generated by ZacrosWrapper\n\nlattice periodic_cell\n\n')
            f.write('cell_vectors       # in row format (Angstroms)\n')
            for row in np.asarray(self.lattice_matrix):
                f.write('\t %.3f \t%.3f \n' % (row[0], row[1]))
            f.write('\nrepeat_cell\t 1 \t 1 \n\n')
            f.write('n_cell_sites \t %d \n' % n)
            f.write('n_site_types \t %d \n' % len(self.site_type_names))
            f.write('site_type_names \t' + ''.join(s + '\t' for s in self.site_type_names) + '\n')
            f.write('site_types \t' + ''.join('%d  ' % t for t in self.site_type_inds) + '\n\n')
            f.write('site_coordinates \t # fractional coordinates (x,y) in row format\n')
            for x, y in (self.frac_coords if n else []):
                f.write('\t %.5f \t%.5f \n' % (x, y))
            f.write('\nneighboring_structure \t # site-neighsite cell\n')
            for (i, j), cell in zip(self.neighbor_list, self.cell_list):
                f.write('\t %d-%d %s \n' % (i + 1, j + 1, cell))
            f.write('end_neighboring_structure\n\nend_lattice\n')
