"""ORACLE support — imports the UNMODIFIED reference (leepaul009/LaneGCN-1) for golden generation.

Only works where the reference tree exists (/root/reference in the authoring container; it does not
exist on the GPU box, so nothing in `-m gpu` tests, smoke() or bench.py calls this).  Nothing is copied:
the reference modules are imported from where they lie, behind the 4-item compatibility shim of
SURVEY §8(c): `fractions.gcd` (removed in py3.9; lanegcn.py:8, layers.py:6), `numpy.bool` (removed in
numpy 1.24; data.py:167,206,521,538), stub modules for argoverse-api / skimage (data.py:11-13), and
`lanegcn.gpu -> identity` for the CPU run (utils.py:84 calls .cuda() unconditionally).
"""
import fractions
import math
import os
import sys
import types

REF_DIR = os.environ.get("LGCN_REFERENCE_DIR", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "lanegcn.py"))


def load():
    """Returns (lanegcn, data) reference modules."""
    import numpy as np

    if not available():
        raise RuntimeError(f"reference tree not found at {REF_DIR}")
    fractions.gcd = math.gcd
    if not hasattr(np, "bool"):
        np.bool = bool
    for name in [
        "argoverse", "argoverse.data_loading", "argoverse.data_loading.argoverse_forecasting_loader",
        "argoverse.map_representation", "argoverse.map_representation.map_api", "skimage", "skimage.transform",
    ]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["argoverse.data_loading.argoverse_forecasting_loader"].ArgoverseForecastingLoader = object
    sys.modules["argoverse.map_representation.map_api"].ArgoverseMap = object
    sys.modules["skimage.transform"].rotate = None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    # the reference imports its siblings by bare name ("data", "utils", "layers"): make sure nothing of
    # ours shadows them, then restore sys.path so those generic names do not leak into later imports.
    import lanegcn as ref_lanegcn  # noqa: E402
    import data as ref_data  # noqa: E402

    ref_lanegcn.gpu = lambda x: x
    return ref_lanegcn, ref_data
