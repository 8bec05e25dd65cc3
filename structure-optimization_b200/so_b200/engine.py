"""Thin object wrappers over the C ABI handles (lattice, surrogate tables, chains).  NumPy in / NumPy out;
all computation happens in libso_b200.so on the GPU."""
import ctypes as C
import numpy as np
from . import _lib as B


def _p(arr, typ):
    return arr.ctypes.data_as(typ) if arr is not None else None


def local_offsets(radius=3):
    """Descriptor column -> (o1, o2) for the (2r+1)^2 window in get_local_inds order (dynamic_cat.py:325-341)."""
    r = range(-radius, radius + 1)
    return np.array([(a, b) for b in r for a in r], dtype=np.int32)


def full_offsets(d):
    """Descriptor column c of a full translation row <-> offset var_ind_to_sym_inds(c) (dynamic_cat.py:233-267)."""
    v = np.arange(d * d)
    return np.stack([v % d, v // d], axis=1).astype(np.int32)


class Lattice(object):
    def __init__(self, d, kind="LSC", device=0):
        self.lib = B.load()
        self.h = C.c_void_p()
        k = B.KIND_NH3 if str(kind).upper().startswith("NH3") else B.KIND_LSC
        B.check(self.lib.so_lattice_create(int(d), k, int(device), C.byref(self.h)))
        vals = [C.c_int() for _ in range(5)]
        B.check(self.lib.so_lattice_info(self.h, *[C.byref(v) for v in vals]))
        self.d, self.N, self.n_nodes, self.nnz, self.W = [v.value for v in vals]
        self.kind, self.device = k, int(device)

    def csr(self):
        rowptr = np.empty(self.n_nodes + 1, dtype=np.int32)
        col = np.empty(self.nnz, dtype=np.int32)
        B.check(self.lib.so_lattice_csr(self.h, _p(rowptr, B.c_i32p), _p(col, B.c_i32p)))
        return rowptr, col

    def positions(self):
        pos = np.empty((self.n_nodes, 3))
        cell = np.empty((2, 2))
        B.check(self.lib.so_lattice_positions(self.h, _p(pos, B.c_f64p), _p(cell, B.c_f64p)))
        return pos, cell

    def close(self):
        if self.h:
            self.lib.so_lattice_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SurrogateTables(object):
    """Device image of a fitted surrogate (weights, scaler, flattened tree)."""

    def __init__(self, lattice, offsets, W1, b1, w2, b2, y_mean=0.0, y_scale=1.0, tree=None, compact=False):
        """compact=True: 24-bit packed first-layer state, 64 bytes per site (so_surrogate_create_compact): window
        descriptors, 17..20 hidden units; so_anneal then runs the pipelined window kernel on 20 % fewer bytes (with a tree:
        so_surrogate_create_compact_gated and the gated window kernel)."""
        self.lib = B.load()
        self.lattice = lattice
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        W1 = np.ascontiguousarray(W1, dtype=np.float64)
        b1 = np.ascontiguousarray(b1, dtype=np.float64)
        w2 = np.ascontiguousarray(np.asarray(w2, dtype=np.float64).reshape(-1))
        K, H = W1.shape
        if offsets.shape != (K, 2) or b1.shape != (H,) or w2.shape != (H,):
            raise ValueError("inconsistent surrogate shapes")
        self.h = C.c_void_p()
        if tree is not None:
            f = np.ascontiguousarray(tree["feature"], dtype=np.int32)
            t = np.ascontiguousarray(tree["thr"], dtype=np.float64)
            l = np.ascontiguousarray(tree["left"], dtype=np.int32)
            r = np.ascontiguousarray(tree["right"], dtype=np.int32)
            a = np.ascontiguousarray(tree["active"], dtype=np.uint8)
            n_nodes = len(f)
        else:
            f = t = l = r = a = None
            n_nodes = 0
        self.compact = bool(compact)
        if compact and tree is not None:
            B.check(self.lib.so_surrogate_create_compact_gated(
                lattice.h, K, H, _p(offsets, B.c_i32p), _p(W1, B.c_f64p), _p(b1, B.c_f64p), _p(w2, B.c_f64p),
                float(np.asarray(b2).reshape(-1)[0]), float(y_mean), float(y_scale), n_nodes, _p(f, B.c_i32p),
                _p(t, B.c_f64p), _p(l, B.c_i32p), _p(r, B.c_i32p), _p(a, B.c_u8p), C.byref(self.h)))
        elif compact:
            B.check(self.lib.so_surrogate_create_compact(
                lattice.h, K, H, _p(offsets, B.c_i32p), _p(W1, B.c_f64p), _p(b1, B.c_f64p), _p(w2, B.c_f64p),
                float(np.asarray(b2).reshape(-1)[0]), float(y_mean), float(y_scale), C.byref(self.h)))
        else:
          B.check(self.lib.so_surrogate_create(
            lattice.h, K, H, _p(offsets, B.c_i32p), _p(W1, B.c_f64p), _p(b1, B.c_f64p), _p(w2, B.c_f64p),
            float(np.asarray(b2).reshape(-1)[0]), float(y_mean), float(y_scale), n_nodes, _p(f, B.c_i32p),
            _p(t, B.c_f64p), _p(l, B.c_i32p), _p(r, B.c_i32p), _p(a, B.c_u8p), C.byref(self.h)))
        vals = [C.c_int() for _ in range(6)]
        B.check(self.lib.so_surrogate_info(self.h, *[C.byref(v) for v in vals]))
        self.K, self.H, self.Hp, self.e_h, self.e_w, self.gated = [v.value for v in vals]
        self.offsets = offsets

    def eval_rows(self, X):
        """(active[n] bool, rate[n] fp64) for explicit descriptor rows X[n, K]."""
        X = np.asarray(X)
        if X.ndim != 2 or X.shape[1] != self.K:
            raise ValueError("descriptor matrix must be [n_rows, %d]" % self.K)
        Xb = np.ascontiguousarray(X != 0, dtype=np.uint8)
        n = Xb.shape[0]
        active = np.empty(n, dtype=np.uint8)
        rate = np.empty(n, dtype=np.float64)
        B.check(self.lib.so_eval_rows(self.h, n, _p(Xb, B.c_u8p), _p(active, B.c_u8p), _p(rate, B.c_f64p)))
        return active.astype(bool), rate

    def score_batch(self, bits, want_lsc_hist=False, want_nh3_exp=False, out=None):
        """rate[n] (+ LSC / NH3-exported histograms) of n bit-packed structures; `out` = (rate, lsc_hist, nh3_exp)
        preallocated (e.g. pinned) arrays."""
        bits = np.ascontiguousarray(bits, dtype=np.uint32)
        n = bits.shape[0]
        if out is not None:
            rate, lh, ne = out
        else:
            rate = np.empty(n, dtype=np.float64)
            lh = np.empty((n, 3), dtype=np.int32) if want_lsc_hist else None
            ne = np.empty((n, 8), dtype=np.int32) if want_nh3_exp else None
        B.check(self.lib.so_score_batch(self.lattice.h, self.h, n, _p(bits, B.c_u32p), _p(rate, B.c_f64p),
                                        _p(lh, B.c_i32p), _p(ne, B.c_i32p)))
        return rate, lh, ne

    def close(self):
        if self.h:
            self.lib.so_surrogate_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pack_bits(occ, W=None):
    """[n, N] 0/1 -> [n, ceil(N/32)] uint32 (site v = bit v&31 of word v>>5)."""
    occ = np.asarray(occ, dtype=np.uint8)
    if occ.ndim == 1:
        occ = occ[None, :]
    n, N = occ.shape
    W = W or (N + 31) // 32
    pad = np.zeros((n, W * 32), dtype=np.uint8)
    pad[:, :N] = occ
    return np.packbits(pad.reshape(n, W, 32), axis=2, bitorder="little").view(np.uint32).reshape(n, W)


def unpack_bits(bits, N):
    bits = np.ascontiguousarray(bits, dtype=np.uint32)
    n = bits.shape[0]
    return np.unpackbits(bits.view(np.uint8).reshape(n, -1), axis=1, bitorder="little")[:, :N]


class Chains(object):
    """n independent structures / SA chains on one device."""

    def __init__(self, lattice, n_chains):
        self.lib = B.load()
        self.lattice = lattice
        self.n = int(n_chains)
        self.h = C.c_void_p()
        self.surrogate = None
        B.check(self.lib.so_chains_create(lattice.h, self.n, C.byref(self.h)))

    # -- occupancy
    def set_occupancy(self, occ, chain0=0):
        occ = np.ascontiguousarray(occ, dtype=np.uint8)
        if occ.ndim == 1:
            occ = occ[None, :]
        if occ.shape[1] != self.lattice.N:
            raise ValueError("occupancy must have %d sites" % self.lattice.N)
        B.check(self.lib.so_set_occupancy(self.h, chain0, occ.shape[0], _p(occ, B.c_u8p)))

    def get_occupancy(self, chain0=0, n=None):
        n = self.n - chain0 if n is None else n
        occ = np.empty((n, self.lattice.N), dtype=np.uint8)
        B.check(self.lib.so_get_occupancy(self.h, chain0, n, _p(occ, B.c_u8p)))
        return occ

    def set_bits(self, bits, chain0=0):
        bits = np.ascontiguousarray(bits, dtype=np.uint32)
        if bits.ndim == 1:
            bits = bits[None, :]
        B.check(self.lib.so_set_occupancy_bits(self.h, chain0, bits.shape[0], _p(bits, B.c_u32p)))

    def get_bits(self, chain0=0, n=None, out=None):
        """Bit-planes of chains [chain0, chain0+n); `out` may be a preallocated (e.g. pinned) uint32 array."""
        n = self.n - chain0 if n is None else n
        bits = np.empty((n, self.lattice.W), dtype=np.uint32) if out is None else out
        if bits.shape != (n, self.lattice.W) or bits.dtype != np.uint32 or not bits.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous uint32 array of shape (%d, %d)" % (n, self.lattice.W))
        B.check(self.lib.so_get_occupancy_bits(self.h, chain0, n, _p(bits, B.c_u32p)))
        return bits

    def randomize(self, seed, coverage=None):
        B.check(self.lib.so_randomize(self.h, int(seed), -1.0 if coverage is None else float(coverage)))

    def island_structures(self, index, n_strucs=48, n_mutate=16, chain0=0):
        """Chains chain0.. become the seeded island structures index[k] of n_strucs
        (drivers/generate_init_structures.py:26-116, same NumPy legacy stream), generated on the device."""
        index = np.ascontiguousarray(index, dtype=np.int32).reshape(-1)
        B.check(self.lib.so_island_structures(self.h, int(chain0), len(index), _p(index, B.c_i32p), int(n_strucs),
                                              int(n_mutate)))

    def flip(self, chains, sites):
        chains = np.ascontiguousarray(chains, dtype=np.int64).reshape(-1)
        sites = np.ascontiguousarray(sites, dtype=np.int32).reshape(-1)
        B.check(self.lib.so_flip(self.h, len(chains), _p(chains, B.c_i64p), _p(sites, B.c_i32p)))

    # -- descriptors
    def descriptors(self, chain0=0, n=None, c6=False, lsc_type=False, nh3_type=False, lsc_hist=False,
                    nh3_hist=False, nh3_exp=False):
        n = self.n - chain0 if n is None else n
        N = self.lattice.N
        out = {}
        if c6: out["c6"] = np.empty((n, N), dtype=np.uint8)
        if lsc_type: out["lsc_type"] = np.empty((n, N), dtype=np.uint8)
        if nh3_type: out["nh3_type"] = np.empty((n, 5 * N), dtype=np.uint8)
        if lsc_hist: out["lsc_hist"] = np.empty((n, 3), dtype=np.int32)
        if nh3_hist: out["nh3_hist"] = np.empty((n, 20), dtype=np.int32)
        if nh3_exp: out["nh3_exp"] = np.empty((n, 8), dtype=np.int32)
        B.check(self.lib.so_descriptors(self.h, chain0, n, _p(out.get("c6"), B.c_u8p), _p(out.get("lsc_type"), B.c_u8p),
                                        _p(out.get("nh3_type"), B.c_u8p), _p(out.get("lsc_hist"), B.c_i32p),
                                        _p(out.get("nh3_hist"), B.c_i32p), _p(out.get("nh3_exp"), B.c_i32p)))
        return out

    def flip_delta(self, sites, chain0=0):
        sites = np.ascontiguousarray(sites, dtype=np.int32).reshape(-1)
        n = len(sites)
        d_lsc = np.empty((n, 3), dtype=np.int32)
        d_nh3 = np.empty((n, 8), dtype=np.int32)
        B.check(self.lib.so_flip_delta(self.h, chain0, n, _p(sites, B.c_i32p), _p(d_lsc, B.c_i32p), _p(d_nh3, B.c_i32p)))
        return d_lsc, d_nh3

    # -- surrogate state
    def attach(self, tables):
        B.check(self.lib.so_chains_attach(self.h, tables.h))
        self.surrogate = tables

    def score(self, chain0=0, n=None):
        n = self.n - chain0 if n is None else n
        of = np.empty(n, dtype=np.float64)
        B.check(self.lib.so_score(self.h, chain0, n, _p(of, B.c_f64p)))
        return of

    def objective(self, chain0=0, n=None, sums=False, out=None):
        n = self.n - chain0 if n is None else n
        of = np.empty(n, dtype=np.float64) if out is None else out
        if not sums:
            B.check(self.lib.so_objective(self.h, chain0, n, _p(of, B.c_f64p), None, None, None))
            return of
        hi = np.empty(n, dtype=np.int64)
        lo = np.empty(n, dtype=np.int64)
        na = np.empty(n, dtype=np.int32)
        B.check(self.lib.so_objective(self.h, chain0, n, _p(of, B.c_f64p), _p(hi, B.c_i64p), _p(lo, B.c_i64p),
                                      _p(na, B.c_i32p)))
        return of, hi, lo, na

    def preact(self, chain):
        hq = np.empty((self.lattice.N, self.surrogate.Hp), dtype=np.int32)
        B.check(self.lib.so_get_preact(self.h, int(chain), _p(hq, B.c_i32p)))
        return hq

    # -- annealing
    def anneal(self, temps, step0=0, seed=0, chain_id0=0, proposals=None, uniforms=None, want_accept=False,
               exact_guard=False, guard_tol=0.0, variant=0):
        """so_anneal: len(temps) Metropolis steps for every chain (the loop of OML/optimize_SA.py:59-132).
        temps[s] is the schedule temperature of step step0 + s (times the chain's tscale); draws come from Philox streams
        keyed by (seed, chain_id0 + chain, step) or from proposals / uniforms [n_chains, steps]; want_accept returns the
        accept/reject bytes.  variant selects a kernel for cross-checks (0 = automatic, see include/so_b200.h)."""
        temps = np.ascontiguousarray(temps, dtype=np.float64)
        ns = len(temps)
        a = B.AnnealArgs()
        a.step0, a.n_steps, a.temps, a.seed, a.chain_id0 = int(step0), ns, _p(temps, B.c_f64p), int(seed), int(chain_id0)
        keep = [temps]
        if proposals is not None:
            proposals = np.ascontiguousarray(proposals, dtype=np.int32).reshape(self.n, ns)
            uniforms = np.ascontiguousarray(uniforms, dtype=np.float32).reshape(self.n, ns)
            a.proposals, a.uniforms = _p(proposals, B.c_i32p), _p(uniforms, B.c_f32p)
            keep += [proposals, uniforms]
        acc = None
        if want_accept:
            acc = np.empty((self.n, ns), dtype=np.uint8)
            a.accept_out = _p(acc, B.c_u8p)
        a.exact_guard, a.guard_tol, a.variant = int(bool(exact_guard)), float(guard_tol), int(variant)
        st = B.AnnealStats()
        B.check(self.lib.so_anneal(self.h, C.byref(a), C.byref(st)))
        stats = dict(flips=st.flips, accepted=st.accepted, guard_fired=st.guard_fired, kernel_ms=st.kernel_ms,
                     kernel=self.lib.so_kernel_name(st.kernel_id).decode())
        return (stats, acc) if want_accept else stats

    def lsc_analytical(self, chain0=0, n=None, site_rates=True):
        """(struct_rate[n], site_rates[n, N] or None): analytical steady state of the LSC model for the current bits."""
        n = self.n - chain0 if n is None else n
        rate = np.empty(n, dtype=np.float64)
        sr = np.empty((n, self.lattice.N), dtype=np.float64) if site_rates else None
        B.check(self.lib.so_lsc_analytical(self.h, chain0, n, _p(sr, B.c_f64p), _p(rate, B.c_f64p)))
        return rate, sr

    def anneal_analytical(self, temps, step0=0, seed=0, chain_id0=0, proposals=None, uniforms=None, want_accept=False):
        temps = np.ascontiguousarray(temps, dtype=np.float64)
        ns = len(temps)
        a = B.AnnealArgs()
        a.step0, a.n_steps, a.temps, a.seed, a.chain_id0 = int(step0), ns, _p(temps, B.c_f64p), int(seed), int(chain_id0)
        if proposals is not None:
            proposals = np.ascontiguousarray(proposals, dtype=np.int32).reshape(self.n, ns)
            uniforms = np.ascontiguousarray(uniforms, dtype=np.float32).reshape(self.n, ns)
            a.proposals, a.uniforms = _p(proposals, B.c_i32p), _p(uniforms, B.c_f32p)
        acc = None
        if want_accept:
            acc = np.empty((self.n, ns), dtype=np.uint8)
            a.accept_out = _p(acc, B.c_u8p)
        of = np.empty(self.n, dtype=np.float64)
        st = B.AnnealStats()
        B.check(self.lib.so_anneal_analytical(self.h, C.byref(a), C.byref(st), _p(of, B.c_f64p)))
        stats = dict(flips=st.flips, accepted=st.accepted, guard_fired=0, kernel_ms=st.kernel_ms, OF=of)
        return (stats, acc) if want_accept else stats

    def set_tscale(self, tscale, chain0=0):
        tscale = np.ascontiguousarray(tscale, dtype=np.float64).reshape(-1)
        B.check(self.lib.so_set_tscale(self.h, chain0, len(tscale), _p(tscale, B.c_f64p)))

    def get_tscale(self, chain0=0, n=None):
        n = self.n - chain0 if n is None else n
        t = np.empty(n, dtype=np.float64)
        B.check(self.lib.so_get_tscale(self.h, chain0, n, _p(t, B.c_f64p)))
        return t

    def track_types(self, enable=True):
        """Carry the LSC / exported-NH3 type histograms of every chain through accepted flips (so_track_types)."""
        B.check(self.lib.so_track_types(self.h, int(bool(enable))))

    def type_histograms(self, chain0=0, n=None):
        """(lsc_hist[n, 3], nh3_exp[n, 8]) of the current structures, as tracked inside the SA kernels (no re-typing)."""
        n = self.n - chain0 if n is None else n
        lh = np.empty((n, 3), dtype=np.int32)
        ne = np.empty((n, 8), dtype=np.int32)
        B.check(self.lib.so_type_histograms(self.h, chain0, n, _p(lh, B.c_i32p), _p(ne, B.c_i32p)))
        return lh, ne

    def save(self, path, step=0, seed=0):
        """Checkpoint (bits, tscale, exact sums) with the caller's step counter and seed; resume is bit-exact."""
        B.check(self.lib.so_chains_save(self.h, str(path).encode(), int(step), int(seed)))

    def load(self, path):
        """Restore a checkpoint written by save() into this handle (same lattice and chain count); returns (step, seed)."""
        step, seed = C.c_int64(), C.c_uint64()
        B.check(self.lib.so_chains_load(self.h, str(path).encode(), C.byref(step), C.byref(seed)))
        return step.value, seed.value

    def best(self):
        of = C.c_double()
        ch = C.c_int64()
        bits = np.empty(self.lattice.W, dtype=np.uint32)
        B.check(self.lib.so_best(self.h, C.byref(of), C.byref(ch), _p(bits, B.c_u32p)))
        return of.value, ch.value, bits

    def close(self):
        if self.h:
            self.lib.so_chains_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Trainer(object):
    """Device-side Adam fit of the K -> H -> 1 ReLU regressor (so_trainer_*; OML/train_surrogate.py:141-152)."""

    def __init__(self, X, y, W1, b1, w2, b2, lr=1e-4, alpha=0.1, beta1=0.9, beta2=0.999, eps=1e-8, tol=1e-5,
                 batch_size=None, n_iter_no_change=10, device=0):
        self.lib = B.load()
        X = np.asarray(X)
        if X.ndim != 2:
            raise ValueError("X must be [n_rows, K]")
        Xu = np.ascontiguousarray(X, dtype=np.uint8)
        if not np.array_equal(Xu, X):
            raise ValueError("training rows must be 0/1 occupancies")
        self.n_rows, self.K = Xu.shape
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1))
        W1 = np.ascontiguousarray(W1, dtype=np.float64)
        self.H = W1.shape[1]
        b1 = np.ascontiguousarray(b1, dtype=np.float64).reshape(-1)
        w2 = np.ascontiguousarray(w2, dtype=np.float64).reshape(-1)
        if y.shape != (self.n_rows,) or W1.shape != (self.K, self.H) or b1.shape != (self.H,) or w2.shape != (self.H,):
            raise ValueError("inconsistent shapes")
        hp = B.TrainParams(float(lr), float(alpha), float(beta1), float(beta2), float(eps), float(tol),
                           int(min(200, self.n_rows) if batch_size is None else np.clip(batch_size, 1, self.n_rows)),
                           int(n_iter_no_change))
        self.batch_size = hp.batch_size
        self.h = C.c_void_p()
        B.check(self.lib.so_trainer_create(int(device), self.n_rows, self.K, self.H, _p(Xu, B.c_u8p), _p(y, B.c_f64p),
                                           _p(W1, B.c_f64p), _p(b1, B.c_f64p), _p(w2, B.c_f64p), float(b2), C.byref(hp),
                                           C.byref(self.h)))

    def run(self, orders):
        """Advance one epoch per row of `orders` [n_epochs, n_rows]; returns (epoch losses actually run, status dict)."""
        orders = np.ascontiguousarray(orders, dtype=np.int32).reshape(-1, self.n_rows)
        loss = np.zeros(len(orders))
        st = B.TrainStatus()
        B.check(self.lib.so_trainer_run(self.h, len(orders), _p(orders, B.c_i32p), _p(loss, B.c_f64p), C.byref(st)))
        return loss[:st.epochs_done], dict(epochs_done=st.epochs_done, n_iter=st.n_iter, stopped=bool(st.stopped),
                                           adam_steps=st.adam_steps, best_loss=st.best_loss, kernel_ms=st.kernel_ms)

    def predict(self, X):
        Xu = np.ascontiguousarray(X, dtype=np.uint8).reshape(-1, self.K)
        if not np.array_equal(Xu, np.asarray(X).reshape(-1, self.K)):
            raise ValueError("rows must be 0/1 occupancies")
        out = np.empty(len(Xu))
        B.check(self.lib.so_trainer_predict(self.h, len(Xu), _p(Xu, B.c_u8p), _p(out, B.c_f64p)))
        return out

    def get(self):
        W1 = np.empty((self.K, self.H))
        b1 = np.empty(self.H)
        w2 = np.empty(self.H)
        b2 = C.c_double()
        B.check(self.lib.so_trainer_get(self.h, _p(W1, B.c_f64p), _p(b1, B.c_f64p), _p(w2, B.c_f64p), C.byref(b2)))
        return W1, b1, w2, b2.value

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.so_trainer_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
