"""ORACLE support — imports the UNMODIFIED reference (leepaul009/LaneGCN-1) for golden generation.

The reference modules are imported from where they lie (/root/reference in the authoring container) or, where that
tree does not exist (the GPU box), from the sourceless byte-compiled build of the same files in oracle/_ref/
(oracle/make_ref.py; git-ignored build output).  No reference source is copied into the repository.  Users: golden
generation and oracle validation in tests/, and bench.py's reference arm (`--impl reference`, `gpu_eager_reference`)
— never the product path.  Everything goes through the 4-item compatibility shim of
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
BUILT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def source_available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "lanegcn.py"))


def built_available() -> bool:
    return os.path.isfile(os.path.join(BUILT_DIR, "lanegcn.refbc"))


class _BuiltFinder:
    """sys.meta_path finder for the byte-compiled reference modules in oracle/_ref (``<module>.refbc`` = CPython
    bytecode; the reference's modules import each other by bare name, so they are resolved here)."""

    NAMES = ("lanegcn", "layers", "utils", "data")

    @classmethod
    def find_spec(cls, name, path=None, target=None):
        import importlib.machinery
        import importlib.util

        f = os.path.join(BUILT_DIR, name + ".refbc")
        if name not in cls.NAMES or not os.path.isfile(f):
            return None
        return importlib.util.spec_from_file_location(name, f, loader=importlib.machinery.SourcelessFileLoader(name, f))


def available() -> bool:
    return source_available() or built_available()


def where() -> str:
    return REF_DIR if source_available() else BUILT_DIR


def load(keep_gpu: bool = False):
    """Returns (lanegcn, data) reference modules.  ``keep_gpu=True`` leaves ``utils.gpu`` intact (the reference's own
    tensor-by-tensor H2D path, for the eager-on-GPU timing); otherwise it is rebound to the identity for CPU runs."""
    import numpy as np

    if not available():
        raise RuntimeError(f"reference not found at {REF_DIR} and no build in {BUILT_DIR} (python oracle/make_ref.py)")
    ref_dir = where()
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
    if source_available():
        if ref_dir not in sys.path:
            sys.path.insert(0, ref_dir)
    elif _BuiltFinder not in sys.meta_path:
        sys.meta_path.insert(0, _BuiltFinder)
    # the reference imports its siblings by bare name ("data", "utils", "layers"): make sure nothing of
    # ours shadows them, then restore sys.path so those generic names do not leak into later imports.
    import lanegcn as ref_lanegcn  # noqa: E402
    import data as ref_data  # noqa: E402

    if not hasattr(ref_lanegcn, "_lgcn_real_gpu"):
        ref_lanegcn._lgcn_real_gpu = ref_lanegcn.gpu
    ref_lanegcn.gpu = ref_lanegcn._lgcn_real_gpu if keep_gpu else (lambda x: x)
    return ref_lanegcn, ref_data
