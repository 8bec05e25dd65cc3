"""ctypes binding of include/so_b200.h.  Loading fails loudly when the CUDA library is missing: there is no
CPU fallback anywhere in this package."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SO_B200_LIB: another build of the same library (same-box A/B measurements of kernel changes)
LIB_PATH = os.environ.get("SO_B200_LIB") or os.path.join(os.path.dirname(_HERE), "libso_b200.so")

SO_OK, SO_EINVAL, SO_ECUDA, SO_ENOMEM, SO_ESTATE = 0, -1, -2, -3, -4
KIND_LSC, KIND_NH3 = 0, 1

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_u8p = C.POINTER(C.c_uint8)
c_u32p = C.POINTER(C.c_uint32)
c_f32p = C.POINTER(C.c_float)
c_f64p = C.POINTER(C.c_double)
c_intp = C.POINTER(C.c_int)


class AnnealArgs(C.Structure):
    _fields_ = [("step0", C.c_int64), ("n_steps", C.c_int64), ("temps", c_f64p), ("seed", C.c_uint64),
                ("chain_id0", C.c_int64), ("proposals", c_i32p), ("uniforms", c_f32p), ("accept_out", c_u8p),
                ("exact_guard", C.c_int), ("guard_tol", C.c_double), ("variant", C.c_int)]


class AnnealStats(C.Structure):
    _fields_ = [("flips", C.c_int64), ("accepted", C.c_int64), ("guard_fired", C.c_int64),
                ("kernel_ms", C.c_float), ("kernel_id", C.c_int)]


class TrainParams(C.Structure):
    _fields_ = [("lr", C.c_double), ("alpha", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double),
                ("eps", C.c_double), ("tol", C.c_double), ("batch_size", C.c_int), ("n_iter_no_change", C.c_int)]


class TrainStatus(C.Structure):
    _fields_ = [("epochs_done", C.c_int), ("n_iter", C.c_int), ("stopped", C.c_int), ("adam_steps", C.c_int64),
                ("best_loss", C.c_double), ("kernel_ms", C.c_float)]


# name -> (restype, argtypes); every symbol declared in include/so_b200.h
PROTOTYPES = {
    "so_version": (C.c_int, []),
    "so_last_error": (C.c_char_p, []),
    "so_device_count": (C.c_int, [c_intp]),
    "so_lattice_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "so_lattice_destroy": (C.c_int, [C.c_void_p]),
    "so_lattice_info": (C.c_int, [C.c_void_p, c_intp, c_intp, c_intp, c_intp, c_intp]),
    "so_lattice_csr": (C.c_int, [C.c_void_p, c_i32p, c_i32p]),
    "so_lattice_positions": (C.c_int, [C.c_void_p, c_f64p, c_f64p]),
    "so_chains_create": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]),
    "so_chains_destroy": (C.c_int, [C.c_void_p]),
    "so_set_occupancy": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, c_u8p]),
    "so_get_occupancy": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, c_u8p]),
    "so_set_occupancy_bits": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, c_u32p]),
    "so_get_occupancy_bits": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, c_u32p]),
    "so_randomize": (C.c_int, [C.c_void_p, C.c_uint64, C.c_double]),
    "so_island_structures": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, c_i32p, C.c_int, C.c_int]),
    "so_flip": (C.c_int, [C.c_void_p, C.c_int64, c_i64p, c_i32p]),
    "so_descriptors": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, c_u8p, c_u8p, c_u8p, c_i32p, c_i32p, c_i32p]),
    "so_flip_delta": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, c_i32p, c_i32p, c_i32p]),
    "so_surrogate_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_i32p, c_f64p, c_f64p, c_f64p, C.c_double,
                                      C.c_double, C.c_double, C.c_int, c_i32p, c_f64p, c_i32p, c_i32p, c_u8p,
                                      C.POINTER(C.c_void_p)]),
    "so_surrogate_destroy": (C.c_int, [C.c_void_p]),
    "so_surrogate_create_compact": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_i32p, c_f64p, c_f64p, c_f64p, C.c_double,
                                              C.c_double, C.c_double, C.POINTER(C.c_void_p)]),
    "so_surrogate_create_compact_gated": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_i32p, c_f64p, c_f64p, c_f64p, C.c_double,
                                                    C.c_double, C.c_double, C.c_int, c_i32p, c_f64p, c_i32p, c_i32p, c_u8p,
                                                    C.POINTER(C.c_void_p)]),
    "so_surrogate_info": (C.c_int, [C.c_void_p, c_intp, c_intp, c_intp, c_intp, c_intp, c_intp]),
    "so_chains_attach": (C.c_int, [C.c_void_p, C.c_void_p]),
    "so_score": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, c_f64p]),
    "so_objective": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, c_f64p, c_i64p, c_i64p, c_i32p]),
    "so_get_preact": (C.c_int, [C.c_void_p, C.c_int64, c_i32p]),
    "so_score_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, c_u32p, c_f64p, c_i32p, c_i32p]),
    "so_eval_rows": (C.c_int, [C.c_void_p, C.c_int64, c_u8p, c_u8p, c_f64p]),
    "so_anneal": (C.c_int, [C.c_void_p, C.POINTER(AnnealArgs), C.POINTER(AnnealStats)]),
    "so_kernel_name": (C.c_char_p, [C.c_int]),
    "so_set_tscale": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, c_f64p]),
    "so_get_tscale": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, c_f64p]),
    "so_track_types": (C.c_int, [C.c_void_p, C.c_int]),
    "so_type_histograms": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, c_i32p, c_i32p]),
    "so_chains_save": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64, C.c_uint64]),
    "so_chains_load": (C.c_int, [C.c_void_p, C.c_char_p, c_i64p, C.POINTER(C.c_uint64)]),
    "so_lsc_analytical": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, c_f64p, c_f64p]),
    "so_anneal_analytical": (C.c_int, [C.c_void_p, C.POINTER(AnnealArgs), C.POINTER(AnnealStats), c_f64p]),
    "so_best": (C.c_int, [C.c_void_p, c_f64p, c_i64p, c_u32p]),
    "so_best_record_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "so_objectives_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "so_exchange": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int64, C.c_double,
                              C.c_void_p, C.c_void_p, C.c_void_p, c_i64p]),
    "so_trainer_create": (C.c_int, [C.c_int, C.c_int64, C.c_int, C.c_int, c_u8p, c_f64p, c_f64p, c_f64p, c_f64p,
                                    C.c_double, C.POINTER(TrainParams), C.POINTER(C.c_void_p)]),
    "so_trainer_destroy": (C.c_int, [C.c_void_p]),
    "so_trainer_run": (C.c_int, [C.c_void_p, C.c_int, c_i32p, c_f64p, C.POINTER(TrainStatus)]),
    "so_trainer_get": (C.c_int, [C.c_void_p, c_f64p, c_f64p, c_f64p, c_f64p]),
    "so_tree_fit": (C.c_int, [C.c_int, C.c_int64, C.c_int, c_u8p, c_u8p, C.c_int, C.c_int, C.c_uint32, C.c_int, c_i32p, c_f64p,
                              c_i32p, c_i32p, c_f64p, c_f64p, c_i32p, C.POINTER(C.c_float)]),
    "so_build_training_set": (C.c_int, [C.c_void_p, C.c_int64, c_u32p, c_f64p, C.c_double, c_u8p, c_u8p, c_f64p]),
    "so_trainer_predict": (C.c_int, [C.c_void_p, C.c_int64, c_u8p, c_f64p]),
}

_lib = None


class SoError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "so_b200 error %d: %s" % (code, msg))
        self.code, self.msg = code, msg


def load():
    """dlopen the in-tree CUDA library and bind every prototype.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s not found: build it with `python structure-optimization_b200/build.py` "
                          "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc):
    """Map a C return code to the reference's exception conventions."""
    if rc == SO_OK:
        return
    msg = load().so_last_error().decode("utf-8", "replace")
    if msg.startswith("Invalid occupancy"):
        raise NameError(msg)                      # OML/dynamic_cat.py:185, OML/optimize_SA.py:83
    if rc == SO_EINVAL:
        raise ValueError(msg)
    if rc == SO_ENOMEM:
        raise MemoryError(msg)
    raise SoError(rc, msg)
