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
