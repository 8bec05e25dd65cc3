"""Build libso_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["csrc/so_api.cu", "csrc/so_desc.cu", "csrc/so_sa.cu", "csrc/so_sa_window.cu", "csrc/so_sa_window_gated.cu", "csrc/so_sa_resident.cu", "csrc/so_lsc.cu", "csrc/so_score_tc.cu", "csrc/so_train.cu", "csrc/so_tree.cu"]
HEADERS = ["csrc/so_internal.cuh", "csrc/so_window_common.cuh", "csrc/so_types.cuh", "../include/so_b200.h"]
LIB = os.path.join(HERE, "libso_b200.so")


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA extension cannot be built")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SOURCES + HEADERS)


OBJ_DIR = os.path.join(HERE, "_obj")          # per-source objects (git-ignored); only changed sources are recompiled


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    flags = ["-Xcompiler", "-fPIC", "-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a"]
    if verbose:
        flags = ["-Xptxas", "-v"] + flags
    hdr_t = max(os.path.getmtime(os.path.join(HERE, h)) for h in HEADERS)

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        src_t = max(os.path.getmtime(os.path.join(HERE, src)), hdr_t)
        if not force and not verbose and os.path.exists(obj) and os.path.getmtime(obj) > src_t:
            return obj, ""
        res = subprocess.run([nvcc, "-c"] + flags + ["-o", obj, src], cwd=HERE, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s%s" % (src, res.stdout, res.stderr))
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    res = subprocess.run([nvcc, "-shared", "-o", LIB] + [o for o, _ in results], cwd=HERE, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write("".join(err for _, err in results))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
