import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "structure-optimization_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    # the CUDA library is a build artefact (git-ignored): build it once if this checkout does not have it yet
    lib = os.path.join(ROOT, "structure-optimization_b200", "libso_b200.so")
    if not os.path.exists(lib):
        spec = importlib.util.spec_from_file_location("so_build", os.path.join(ROOT, "structure-optimization_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
