"""`OML` alias: lets reference-style driver code (`from OML.LSC_cat import *`, `from OML.optimize_SA import *`, and the
module names the upstream drivers actually import, OML.NiPt_NH3 / OML.toy_cat) run on the B200 package unchanged."""
import sys
import so_b200
from so_b200 import dynamic_cat as _dc, LSC_cat as _lsc, NH3_cat as _nh3, optimize_SA as _opt, train_surrogate as _ts
from so_b200 import *                                         # noqa: F401,F403

for _name, _mod in (("dynamic_cat", sys.modules["so_b200.dynamic_cat"]), ("LSC_cat", sys.modules["so_b200.LSC_cat"]),
                    ("NH3_cat", sys.modules["so_b200.NH3_cat"]), ("optimize_SA", sys.modules["so_b200.optimize_SA"]),
                    ("train_surrogate", sys.modules["so_b200.train_surrogate"]),
                    ("NiPt_NH3", sys.modules["so_b200.NH3_cat"]), ("toy_cat", sys.modules["so_b200.LSC_cat"])):
    sys.modules["OML." + _name] = _mod
