"""ORACLE (test infrastructure, NOT product code) -- ctypes front end of oracle/so_oracle.c (many chains, OpenMP)."""
import ctypes as C
import numpy as np
from . import build_oracle
from . import surrogate as S

_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle.build())
        _lib.oracle_anneal_fixed.restype = C.c_int
        _lib.oracle_anneal_fp64.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def anneal_fixed_many(p, q, d, x0, temps, prop, u, tscale=None):
    """Same contract as oracle.sa.anneal_fixed, for [n_chains, ...] inputs.  Returns dict of arrays."""
    lib = _load()
    x = np.ascontiguousarray(x0, dtype=np.uint8).copy()
    n_chains, N = x.shape
    temps = np.ascontiguousarray(temps, dtype=np.float64)
    n_steps = len(temps)
    prop = np.ascontiguousarray(prop, dtype=np.int32).reshape(n_chains, n_steps)
    u = np.ascontiguousarray(u, dtype=np.float32).reshape(n_chains, n_steps)
    ts = None if tscale is None else np.ascontiguousarray(tscale, dtype=np.float64)
    off = np.ascontiguousarray(p.offsets, dtype=np.int32)
    W1q = np.ascontiguousarray(q["W1q"], dtype=np.int32)
    b1q = np.ascontiguousarray(q["b1q"], dtype=np.int32)
    w2q = np.ascontiguousarray(q["w2q"], dtype=np.int32)
    t = p.tree
    if t is not None:
        tf = np.ascontiguousarray(t["feature"], dtype=np.int32)
        tt = np.ascontiguousarray(t["thr"], dtype=np.float64)
        tl = np.ascontiguousarray(t["left"], dtype=np.int32)
        tr = np.ascontiguousarray(t["right"], dtype=np.int32)
        ta = np.ascontiguousarray(t["active"], dtype=np.uint8)
        n_tree = len(tf)
    else:
        tf = tt = tl = tr = ta = None
        n_tree = 0
    acc = np.empty((n_chains, n_steps), dtype=np.uint8)
    hi = np.empty(n_chains, dtype=np.int64)
    lo = np.empty(n_chains, dtype=np.int64)
    na = np.empty(n_chains, dtype=np.int32)
    of = np.empty(n_chains, dtype=np.float64)
    rc = lib.oracle_anneal_fixed(
        C.c_int(d), C.c_int(p.K), C.c_int(q["Hp"]), _p(off, C.c_int32), _p(W1q, C.c_int32), _p(b1q, C.c_int32),
        _p(w2q, C.c_int32), C.c_double(q["c"]), C.c_double(p.b2), C.c_double(p.y_scale), C.c_double(p.y_mean),
        C.c_int(n_tree), _p(tf, C.c_int32), _p(tt, C.c_double), _p(tl, C.c_int32), _p(tr, C.c_int32), _p(ta, C.c_uint8),
        C.c_int64(n_chains), C.c_int64(n_steps), _p(x, C.c_uint8), _p(temps, C.c_double), _p(ts, C.c_double),
        _p(prop, C.c_int32), _p(u, C.c_float), _p(acc, C.c_uint8), _p(hi, C.c_int64), _p(lo, C.c_int64),
        _p(na, C.c_int32), _p(of, C.c_double))
    if rc != 0:
        raise MemoryError("oracle_anneal_fixed failed")
    return dict(x=x, accept=acc.astype(bool), hi=hi, lo=lo, n_act=na, OF=of)


def anneal_fp64_many(p, d, x0, temps, prop, u, tscale=None, margin_thr=1e-9):
    """The reference's fp64 decision rule with the ORIGINAL weights (oracle_anneal_fp64) for [n_chains, ...] inputs.
    Returns dict(x, accept, min_margin[n_chains], n_below[n_chains], OF)."""
    lib = _load()
    x = np.ascontiguousarray(x0, dtype=np.uint8).copy()
    n_chains, N = x.shape
    temps = np.ascontiguousarray(temps, dtype=np.float64)
    n_steps = len(temps)
    prop = np.ascontiguousarray(prop, dtype=np.int32).reshape(n_chains, n_steps)
    u = np.ascontiguousarray(u, dtype=np.float32).reshape(n_chains, n_steps)
    ts = None if tscale is None else np.ascontiguousarray(tscale, dtype=np.float64)
    off = np.ascontiguousarray(p.offsets, dtype=np.int32)
    W1 = np.ascontiguousarray(p.W1, dtype=np.float64)
    b1 = np.ascontiguousarray(p.b1, dtype=np.float64)
    w2 = np.ascontiguousarray(p.w2, dtype=np.float64)
    t = p.tree
    if t is not None:
        tf = np.ascontiguousarray(t["feature"], dtype=np.int32)
        tt = np.ascontiguousarray(t["thr"], dtype=np.float64)
        tl = np.ascontiguousarray(t["left"], dtype=np.int32)
        tr = np.ascontiguousarray(t["right"], dtype=np.int32)
        ta = np.ascontiguousarray(t["active"], dtype=np.uint8)
        n_tree = len(tf)
    else:
        tf = tt = tl = tr = ta = None
        n_tree = 0
    acc = np.empty((n_chains, n_steps), dtype=np.uint8)
    mm = np.empty(n_chains, dtype=np.float64)
    nb = np.empty(n_chains, dtype=np.int64)
    of = np.empty(n_chains, dtype=np.float64)
    rc = lib.oracle_anneal_fp64(
        C.c_int(d), C.c_int(p.K), C.c_int(p.H), _p(off, C.c_int32), _p(W1, C.c_double), _p(b1, C.c_double),
        _p(w2, C.c_double), C.c_double(p.b2), C.c_double(p.y_scale), C.c_double(p.y_mean), C.c_int(n_tree),
        _p(tf, C.c_int32), _p(tt, C.c_double), _p(tl, C.c_int32), _p(tr, C.c_int32), _p(ta, C.c_uint8),
        C.c_int64(n_chains), C.c_int64(n_steps), _p(x, C.c_uint8), _p(temps, C.c_double), _p(ts, C.c_double),
        _p(prop, C.c_int32), _p(u, C.c_float), C.c_double(margin_thr), _p(acc, C.c_uint8), _p(mm, C.c_double),
        _p(nb, C.c_int64), _p(of, C.c_double))
    if rc != 0:
        raise MemoryError("oracle_anneal_fp64 failed")
    return dict(x=x, accept=acc.astype(bool), min_margin=mm, n_below=nb, OF=of)
