"""Import alias: ``import lanegcn_b200`` loads the package that lives in ``lanegcn-1_b200/``.

The directory name carries a hyphen (it is the name the build contract fixes), which Python cannot
import directly, so this one-file shim loads it under an importable name and then replaces itself in
``sys.modules``.  ``import lanegcn_b200.lanegcn`` etc. resolve through ``submodule_search_locations``.
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lanegcn-1_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
