"""Import helper: the package directory is named `rnb-neus2_b200` (hyphen), which is not a Python identifier."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))


def load_package():
    name = "rnb_neus2_b200"
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(_ROOT, "rnb-neus2_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_scene():
    load_package()
    import importlib
    return importlib.import_module("rnb_neus2_b200.scene")
