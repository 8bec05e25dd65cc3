"""
LSC_cat -- drop-in for OML/LSC_cat.py:13-316 ("linear site coupling" toy catalyst).  The NetworkX graph and the VF2
site typing are replaced by the engine's CSR tables and the descriptor kernel (so_descriptors); the public
attributes (`defected_graph`, `KMC_lat`, `var_ind_kmc_sites`) keep their meaning.
"""
import numpy as np
from .dynamic_cat import dynamic_cat, SlabTemplate
from .graph import GraphView, KMCLattice


class _GraphCat(dynamic_cat):
    """Shared part of LSC_cat / NH3_cat: 5-layer slab, layer 3 variable, graph view (OML/LSC_cat.py:19-112)."""

    def __init__(self, dimen, device=0):
        dynamic_cat.__init__(self, dim1=dimen, dim2=dimen, fixed_layers=4, variable_layers=1, device=device)
        self.KMC_lat = None
        n = self.atoms_per_layer
        symbols = self.atoms_template.get_chemical_symbols()
        for i in range(3 * n, 4 * n):
            symbols[i] = 'Ni'
        for i in range(4 * n, 5 * n):
            symbols[i] = 'Cu'                       # top layer is built as Cu and treated as 'vacuum' in the graph
        pos, cell = self._lattice.positions()
        self.atoms_template = SlabTemplate(pos, cell, symbols)
        self.variable_atoms = range(3 * n, 4 * n)   # 2nd layer from the top is variable (OML/LSC_cat.py:49)
        self.variable_occs = [1 for i in self.variable_atoms]
        rowptr, col = self._lattice.csr()
        self.defected_graph = GraphView(self, rowptr, col)
        self.occs_to_atoms()
        self.occs_to_graph()

    def occs_to_graph(self):
        """Elements are derived from `variable_occs` on access; validate the vector like upstream would fail."""
        self._occs_array()

    def graph_to_occs(self):
        self._occs_array()

    def _new_kmc_lattice(self, names):
        lat = KMCLattice()
        lat.lattice_matrix = self.atoms_template.get_cell()[0:2, 0:2]
        lat.site_type_names = names
        return lat


class LSC_cat(_GraphCat):
    _kind = "LSC"

    def __init__(self, dimen=12, device=0):
        _GraphCat.__init__(self, dimen, device=device)

    def graph_to_KMClattice(self):
        """OML/LSC_cat.py:164-245: types 1 (Ni with 6 Ni neighbours) and 2 (Ni, 4 Ni neighbours, two adjacent
        vacancies), exported in ascending atom order."""
        if self.defected_graph is None:
            raise NameError('Lattice graph not yet defined.')
        n = self.atoms_per_layer
        types = self._push().descriptors(lsc_type=True)["lsc_type"][0]
        self.defected_graph.site_type[:] = 0
        self.defected_graph.site_type[3 * n:4 * n] = types
        self.KMC_lat = self._new_kmc_lattice(['top_Ni', 'top_edge_Ni'])
        sites = np.nonzero(types)[0]
        self.KMC_lat.site_type_inds = [int(t) for t in types[sites]]
        self.var_ind_kmc_sites = [int(v) for v in sites]
        if len(sites) > 0:
            pos = self.atoms_template.get_positions()
            self.KMC_lat.set_cart_coords(pos[3 * n + sites, 0:2])
            self.KMC_lat.Build_neighbor_list(cut=2.77 + 0.1)

    graph_to_KMClattice_V2 = graph_to_KMClattice      # the unfinished upstream variant (OML/LSC_cat.py:114-162)

    def eval_struc_rate_anal(self, x):
        """OML/LSC_cat.py:248-257: analytical structure rate (not normalized) of occupancy x."""
        self.assign_occs(x)
        self.graph_to_KMClattice()
        return np.sum(self.compute_site_rates_lsc())

    def compute_site_rates_lsc(self):
        """OML/LSC_cat.py:260-316 (all rate constants 1): typing, assembly of the steady-state system over the typed
        sites and the dense solve run in one kernel (so_lsc_analytical); rate = coverage on the edge sites."""
        _, site_rates = self._push().lsc_analytical(0, 1)
        return site_rates[0]
