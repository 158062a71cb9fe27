"""Test-harness glue: import the *real* reference from /root/reference (read-only, only present in
the build container) with the three shims SURVEY.md §8c lists.  Nothing here is product code and
nothing on the GPU box may need it: callers must skip when ``load_reference()`` returns None."""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"
_cache = {}


def load_reference():
    if "mods" in _cache:
        return _cache["mods"]
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "NJODE")):
        _cache["mods"] = None
        return None
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].colors = sys.modules["matplotlib.colors"]
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "infty"):
        np.infty = np.inf
    import NJODE.models as ref_models
    import NJODE.data_utils as ref_data_utils
    import NJODE.stock_model as ref_stock_model
    _cache["mods"] = types.SimpleNamespace(models=ref_models, data_utils=ref_data_utils,
                                           stock_model=ref_stock_model)
    return _cache["mods"]
