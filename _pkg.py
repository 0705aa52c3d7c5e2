"""Imports the package directory `vectorchord-bm25_b200/` (hyphenated, as the project is named) under the
importable module name `vectorchord_bm25_b200`."""
import importlib.util
import os
import sys

_NAME = "vectorchord_bm25_b200"


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    root = os.path.dirname(os.path.abspath(__file__))
    path = os.path.join(root, "vectorchord-bm25_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location(_NAME, path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
