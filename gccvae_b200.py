"""Import shim: exposes the package directory `semi-supervised-gated-lt-vae_b200/` (whose name is
not a valid Python identifier) as the module `gccvae_b200`, submodules included."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "semi-supervised-gated-lt-vae_b200")
_spec = importlib.util.spec_from_file_location(
    "gccvae_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["gccvae_b200"] = _mod
_spec.loader.exec_module(_mod)
