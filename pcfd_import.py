"""Registers the in-tree package directory `porous-cfd_b200/` under the importable name
`porous_cfd_b200` (a hyphen cannot appear in a Python module name).

    import pcfd_import; pcfd = pcfd_import.load()
    from porous_cfd_b200.models.pipn.pipn_foam import PipnFoamPp
"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, 'porous-cfd_b200')
NAME = 'porous_cfd_b200'


def load():
    if NAME in sys.modules:
        return sys.modules[NAME]
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    spec = importlib.util.spec_from_file_location(NAME, os.path.join(PKG_DIR, '__init__.py'),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    spec.loader.exec_module(mod)
    return mod
