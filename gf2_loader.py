"""Import helper: the package directory is named `ground-fusion2_b200` (hyphen, per the repo contract), which
Python cannot import by name. `load()` registers it as module `gf2_b200`."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "ground-fusion2_b200")


def load():
    if "gf2_b200" in sys.modules:
        return sys.modules["gf2_b200"]
    spec = importlib.util.spec_from_file_location(
        "gf2_b200", os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["gf2_b200"] = mod
    spec.loader.exec_module(mod)
    return mod
