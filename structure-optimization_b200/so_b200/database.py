"""
The data formats either side of the SA path, so that a structure found on the GPU can be handed to
the external KMC evaluator exactly as the reference does, and a database written by the reference can be read back:

  write_structure_files   OML/KMC_handler.py:13-55   occupancies.npy, occ_symmetries.npy, lattice_input.dat
  read_database           drivers/Online_learn_main.py:26-47
  make_random_structure   drivers/generate_init_structures.py:26-124  (seeded 19-site-island structures)

Pictures (`cat.show`, `KMC_lat.PlotLattice`) need ASE / matplotlib and are skipped; Zacros itself stays external.
"""
import os
import shutil
import numpy as np

ISLAND = [(-2, 2), (-1, 2), (0, 2), (-2, 1), (-1, 1), (0, 1), (1, 1), (-2, 0), (-1, 0), (0, 0), (1, 0), (2, 0),
          (-1, -1), (0, -1), (1, -1), (2, -1), (0, -2), (1, -2), (2, -2)]


def write_structure_files(cat, run_folder, all_symmetries=None, clear_fldr=False):
    if not os.path.exists(run_folder):
        os.makedirs(run_folder)
    if clear_fldr:
        for name in os.listdir(run_folder):
            path = os.path.join(run_folder, name)
            try:
                if os.path.isfile(path):
                    os.unlink(path)
                elif os.path.isdir(path):
                    shutil.rmtree(path)
            except Exception as e:
                print(e)
    np.save(os.path.join(run_folder, 'occupancies.npy'), cat.variable_occs)
    if all_symmetries is None:
        all_symmetries = cat.generate_all_translations_and_rotations()
    np.save(os.path.join(run_folder, 'occ_symmetries.npy'), all_symmetries)
    if cat.KMC_lat is None:
        cat.graph_to_KMClattice()
    cat.KMC_lat.Write_lattice_input(run_folder)


def read_database(fldr_list):
    """[structure_occs (all symmetries stacked), site_rates] of the finished structures."""
    sym_list, site_rate_list = [], []
    for fldr in fldr_list:
        if os.path.isfile(os.path.join(fldr, 'site_rates.npy')):        # skip incomplete calculations
            sym_list.append(np.load(os.path.join(fldr, 'occ_symmetries.npy')))
            site_rate_list.append(np.load(os.path.join(fldr, 'site_rates.npy')))
    return [np.vstack(sym_list), np.vstack(site_rate_list)]


def make_random_structures_device(lattice, indices, n_strucs, n_mutate=16):
    """The same recipe for many structures at once, generated on the device (so_island_structures): returns the
    occupancies [len(indices), N] uint8; bit-equal to make_random_structure(i, n_strucs, ...) for every i."""
    from . import engine
    indices = np.asarray(indices, dtype=np.int32).reshape(-1)
    ch = engine.Chains(lattice, len(indices))
    try:
        ch.island_structures(indices, n_strucs=n_strucs, n_mutate=n_mutate)
        return ch.get_occupancy()
    finally:
        ch.close()


def make_random_structure(i, n_strucs, cat, fldr_name=None):
    """Seeded island structure number i of n_strucs: n_tops 19-site hexagonal islands, then 16 random mutations."""
    np.random.seed(i)
    cat.variable_occs = [0 for j in range(cat.atoms_per_layer)]
    n_tops = int((3 + 15 * float(i) / (n_strucs + 1)))
    top_sites = np.random.choice(range(len(cat.variable_atoms)), size=n_tops, replace=False)
    for t in top_sites:
        s1, s2 = cat.var_ind_to_sym_inds(t)
        for (a, b) in ISLAND:
            cat.variable_occs[cat.sym_inds_to_var_ind(s1 + a, s2 + b)] = 1
    for site in np.random.choice(range(len(cat.variable_atoms)), size=16, replace=False):
        cat.variable_occs[site] = 1 - cat.variable_occs[site]
    cat.occs_to_atoms()
    cat.occs_to_graph()
    cat.graph_to_KMClattice()
    if fldr_name is not None:
        write_structure_files(cat, fldr_name)
    return cat.variable_occs
