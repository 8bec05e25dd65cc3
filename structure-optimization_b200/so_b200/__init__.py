"""B200-native drop-in for the surrogate-driven simulated-annealing path of VlachosGroup/Structure-Optimization.

    from so_b200 import LSC_cat, NH3_cat, surrogate, optimize

mirrors `from OML.LSC_cat import *` etc. of the reference.  Importing the package loads the CUDA library
(libso_b200.so, built by structure-optimization_b200/build.py); there is no CPU fallback.
"""
from . import _lib

_lib.load()

from .dynamic_cat import dynamic_cat, dyn_cat_dist          # noqa: E402
from .LSC_cat import LSC_cat                                # noqa: E402
from .NH3_cat import NH3_cat                                # noqa: E402
from .train_surrogate import surrogate                      # noqa: E402
from .optimize_SA import optimize                           # noqa: E402

__all__ = ["dynamic_cat", "dyn_cat_dist", "LSC_cat", "NH3_cat", "surrogate", "optimize"]
