"""Loader for the hyphenated package directory gmx-acqueduct_b200/."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "gmx-acqueduct_b200")


def load_capi():
    """Import gmx-acqueduct_b200/capi.py as module `gmx_acqueduct_b200_capi`."""
    name = "gmx_acqueduct_b200_capi"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(PKG_DIR, "capi.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod
