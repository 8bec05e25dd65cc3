"""CPU: the C-ABI library loads and exports every symbol include/so_b200.h declares (no compute without a GPU);
host-side logic of the Python layer that needs no device."""
import ctypes
import os
import re
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "so_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(so_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from so_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libso_b200.so does not export %s" % n
        assert n in _lib.PROTOTYPES, "no ctypes prototype for %s" % n
    assert sorted(_lib.PROTOTYPES) == names
    assert lib.so_version() >= 100


def test_error_convention_without_gpu():
    """Argument errors are reported through the return code + so_last_error(), never by crashing."""
    from so_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.so_lattice_create(2, 0, 0, ctypes.byref(h))
    assert rc == _lib.SO_EINVAL and b"d=2" in lib.so_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)
    assert lib.so_lattice_info(None, None, None, None, None, None) == _lib.SO_EINVAL


def test_missing_library_fails_loudly(monkeypatch):
    from so_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libso_b200.so")
    with pytest.raises(ImportError):
        _lib.load()


def test_bit_packing_roundtrip():
    from so_b200 import engine
    x = (np.random.default_rng(0).random((5, 144)) < 0.5).astype(np.uint8)
    bits = engine.pack_bits(x)
    assert bits.shape == (5, 5) and bits.dtype == np.uint32
    assert ((bits[:, 4] >> 16) == 0).all()                              # bits beyond site 143 stay zero
    assert (engine.unpack_bits(bits, 144) == x).all()
    v = 77
    assert ((bits[2, v >> 5] >> (v & 31)) & 1) == x[2, v]
    assert (engine.local_offsets(3)[:3] == [[-3, -3], [-2, -3], [-1, -3]]).all() and len(engine.full_offsets(12)) == 144


def test_cooling_schedule_matches_reference_formulas():
    from so_b200.optimize_SA import cooling_schedule
    from oracle import sa as SA
    for c in ("exp", "log", "linear", "quench"):
        np.testing.assert_array_equal(cooling_schedule(c, 0.05, 777), SA.schedule(c, 0.05, 777))
    np.testing.assert_array_equal(cooling_schedule("linear", 0.05, 50, py2_linear=True), SA.schedule("linear", 0.05, 50, True))
    with pytest.raises(NameError):
        cooling_schedule("nope", 1.0, 3)


def test_descriptor_offsets_and_tree_flattening():
    from so_b200.train_surrogate import descriptor_offsets
    from oracle import lattice as L
    assert (descriptor_offsets(144, 12) == L.full_offsets(12)).all()
    assert (descriptor_offsets(49, 96) == L.local_offsets(3)).all()
    with pytest.raises(ValueError):
        descriptor_offsets(50, 12)


def test_device_regressor_host_logic_with_stub_trainer(monkeypatch):
    """DeviceMLPRegressor's host side without a GPU: the Glorot draws, the cumulative epoch shuffles prepared in chunks by
    a helper thread, and the rewind of the generator when the stop rule fires inside a chunk must reproduce
    scikit-learn's use of the RandomState.  The device trainer is replaced by the NumPy training oracle."""
    import warnings
    from sklearn.neural_network import MLPRegressor
    from so_b200 import engine
    from so_b200.train_surrogate import DeviceMLPRegressor
    from oracle import train as T

    class StubTrainer(object):
        def __init__(self, X, y, W1, b1, w2, b2, lr, alpha, beta1, beta2, eps, tol, batch_size, n_iter_no_change, device):
            self.X, self.y, self.kw = np.asarray(X, dtype=np.float64), y, dict(lr=lr, alpha=alpha, tol=tol)
            self.init, self.orders, self.res = (W1, b1, w2, b2), [], None

        def run(self, orders):
            before = len(self.orders)
            self.orders.extend(list(orders))
            self.res = T.fit_adam(self.X, self.y, init=self.init, orders=self.orders, max_iter=len(self.orders), **self.kw)
            done = self.res["n_iter"] - before
            stopped = self.res["n_iter"] < len(self.orders) or self._stop_at_end()
            return self.res["loss_curve"][before:], dict(epochs_done=done, n_iter=self.res["n_iter"], stopped=stopped,
                                                         adam_steps=0, best_loss=float(self.res["loss_curve"].min()),
                                                         kernel_ms=0.0)

        def _stop_at_end(self):      # did the tol rule fire exactly at the last prepared epoch?
            again = T.fit_adam(self.X, self.y, init=self.init, orders=self.orders + [self.orders[-1]],
                               max_iter=len(self.orders) + 1, **self.kw)
            return again["n_iter"] == len(self.orders)

        def get(self):
            return self.res["W1"], self.res["b1"], self.res["w2"], self.res["b2"]

        def close(self):
            pass

    monkeypatch.setattr(engine, "Trainer", StubTrainer)
    rng = np.random.default_rng(5)
    X = (rng.random((400, 30)) < 0.5).astype(np.float64)
    y = np.maximum(X @ (rng.normal(size=30) / np.sqrt(30)), 0) + 0.05 * rng.normal(size=400)
    y = (y - y.mean()) / y.std()
    kw = dict(activation='relu', learning_rate_init=0.01, alpha=0.1, hidden_layer_sizes=(20,), max_iter=2000, tol=1e-3)
    rs_ref, rs_dev = np.random.RandomState(3), np.random.RandomState(3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = MLPRegressor(random_state=rs_ref, **kw).fit(X, y)
        dev = DeviceMLPRegressor(random_state=rs_dev, epochs_per_launch=7, **kw).fit(X, y)
    assert dev.n_iter_ == ref.n_iter_ and ref.n_iter_ % 7 != 0
    np.testing.assert_allclose(dev.loss_curve_, ref.loss_curve_, rtol=1e-9)
    np.testing.assert_allclose(dev.coefs_[0], ref.coefs_[0], rtol=1e-7, atol=1e-11)
    assert rs_ref.randint(1 << 30) == rs_dev.randint(1 << 30)        # generator left where scikit-learn leaves it


def test_lattice_input_writer_regenerates_reference_fixture(tmp_path):
    """Write_lattice_input byte for byte: the reference's zacros_input/lattice_input.dat rebuilt from its own 14 sites
    (layer-3 atoms of the 12x12 slab) -- compared with the file itself where the reference tree is present, and with its
    SHA-256 everywhere (the file is not copied into this repo)."""
    import hashlib
    import json
    from so_b200.graph import KMCLattice
    from oracle import lattice as L
    import helpers as Hp
    d = 12
    sites = [(5, 1), (11, 1), (0, 2), (4, 2), (1, 7), (8, 7), (2, 8), (8, 8), (3, 9), (6, 9), (1, 10), (2, 11), (7, 11), (10, 11)]
    pos = L.slab_positions(d)[3 * d * d:4 * d * d, :2]
    lat = KMCLattice()
    lat.lattice_matrix = L.cell_vectors(d)
    lat.site_type_names = ['top_Ni', 'top_corner_Ni', 'top_edge_Ni']
    lat.site_type_inds = [2] * len(sites)
    lat.set_cart_coords([pos[i2 * d + i1] for i1, i2 in sites])
    lat.Build_neighbor_list(cut=2.77 + 0.1)
    fx = json.load(open(os.path.join(Hp.GOLDEN, "lattice_input_fixture.json")))
    assert [[i + 1, j + 1] for i, j in lat.neighbor_list] == fx["neighbor_pairs"] and lat.cell_list == ['self', 'self']
    lat.Write_lattice_input(str(tmp_path))
    got = open(os.path.join(str(tmp_path), 'lattice_input.dat'), 'rb').read()
    assert hashlib.sha256(got).hexdigest() == '8d51532be1b424aba1920d194a34fa7057137c6352ad2d1ff03a37626bee3ae8'
    ref = '/root/reference/zacros_input/lattice_input.dat'
    if os.path.exists(ref):
        assert got == open(ref, 'rb').read()


def test_kmc_neighbour_cells_across_the_periodic_edge():
    """Pairs that cross the cell edge carry Zacros' direction keyword (second site displaced into that image)."""
    from so_b200.graph import KMCLattice
    from oracle import lattice as L
    d = 6
    pos = L.slab_positions(d)[3 * d * d:4 * d * d, :2]
    lat = KMCLattice()
    lat.lattice_matrix = L.cell_vectors(d)
    lat.site_type_names = ['top_Ni', 'top_edge_Ni']
    lat.site_type_inds = [1] * (d * d)
    lat.set_cart_coords(pos)
    lat.Build_neighbor_list(cut=2.77 + 0.1)
    assert len(lat.neighbor_list) == 3 * d * d                     # every site has 6 neighbours, each pair once
    a, b = lat.lattice_matrix
    shift = dict(self=0 * a, north=b, northeast=a + b, east=a, southeast=a - b)
    seen = set()
    for (i, j), cell in zip(lat.neighbor_list, lat.cell_list):
        assert abs(np.linalg.norm(pos[j] + shift[cell] - pos[i]) - L.D_NN) < 1e-9
        assert cell != 'self' or i < j
        seen.add(frozenset((i, j)))
    assert len(seen) == 3 * d * d
    assert set(lat.cell_list) == {'self', 'north', 'east', 'southeast'} or 'northeast' in lat.cell_list
    v = lambda i1, i2: (i2 % d) * d + i1 % d
    k = lat.neighbor_list.index([v(5, 2), v(0, 2)])                # (5,2) -> (6,2) wraps along the first cell vector
    assert lat.cell_list[k] == 'east'
    k = lat.neighbor_list.index([v(3, 5), v(3, 0)])                # (3,5) -> (3,6) wraps along the second
    assert lat.cell_list[k] == 'north'
