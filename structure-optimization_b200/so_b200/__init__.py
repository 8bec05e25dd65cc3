"""B200-native drop-in for the surrogate-driven simulated-annealing path of VlachosGroup/Structure-Optimization.
Importing the package loads the CUDA library (libso_b200.so); there is no CPU fallback."""
from . import _lib

_lib.load()
