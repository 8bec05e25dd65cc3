"""
NH3_cat -- drop-in for OML/NH3_cat.py:13-428 (NH3-decomposition catalyst, 19 site types over layers 1-4).
Site typing runs in the descriptor kernel (closed forms of the 11 VF2 matches); the upstream
`flip_atom` (OML/NH3_cat.py:115-155) references undefined names -- here it is the base flip, and the incremental
re-typing it was meant to provide is `flip_delta()`.
"""
import numpy as np
from .LSC_cat import _GraphCat

SITE_IND_DICT = {1: 1, 3: 2, 5: 3, 16: 4, 17: 5, 18: 6, 19: 7}           # OML/NH3_cat.py:421


class NH3_cat(_GraphCat):
    _kind = "NH3"

    def __init__(self, dimen=16, device=0):
        _GraphCat.__init__(self, dimen, device=device)

    def graph_to_KMClattice(self):
        """OML/NH3_cat.py:158-428: all 19 types into `defected_graph`, the 7 exported ones into `KMC_lat`."""
        if self.defected_graph is None:
            raise NameError('Lattice graph not yet defined.')
        types = self._push().descriptors(nh3_type=True)["nh3_type"][0]
        self.defected_graph.site_type[:] = types
        self.KMC_lat = self._new_kmc_lattice(['fcc_Ni', 'top_Ni', 'top_edge_Ni', 'fcc_edge_Ni_3fcc', 'fcc_edge_Ni_3hcp',
                                              'hcp_edge_Ni_3fcc', 'hcp_edge_Ni_3hcp'])
        lut = np.zeros(20, dtype=np.int64)
        for k, v in SITE_IND_DICT.items():
            lut[k] = v
        codes = lut[types]
        nodes = np.nonzero(codes)[0]
        self.KMC_lat.site_type_inds = [int(c) for c in codes[nodes]]
        self.kmc_site_atoms = [int(i) for i in nodes]
        pos = self.atoms_template.get_positions()
        self.KMC_lat.set_cart_coords(pos[nodes, 0:2])
        self.KMC_lat.Build_neighbor_list(cut=2.77 + 0.1)

    def site_type_histogram(self):
        """Counts of the 19 types (index 0 = None) and of the 7 exported KMC types (index 0 = not exported)."""
        out = self._push().descriptors(nh3_hist=True, nh3_exp=True)
        return out["nh3_hist"][0], out["nh3_exp"][0]

    def flip_delta(self, ind):
        """Change of the exported-type histogram if atom `ind` were flipped (nothing is applied)."""
        site = self.variable_atoms.index(ind)
        _, d_nh3 = self._push().flip_delta(np.array([site], dtype=np.int32))
        return d_nh3[0]
