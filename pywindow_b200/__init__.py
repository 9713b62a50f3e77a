"""pywindow_b200 -- B200 (sm_100a) implementation of pywindow's per-frame pore-geometry path.

Drop-in for the reference's public names on that path (``pywindow/__init__.py:3-19``): ``DLPOLY``,
``PDB``, ``XYZ``, ``Input``, ``MolecularSystem``, ``Molecule``, ``compare_properties_dict``,
``make_supercell``, ``periodic_table``.  See DESIGN.md / INTEGRATION.md.
"""
_LAZY = {
    "Engine": "engine", "default_engine": "engine",
    "MolecularSystem": "molecular", "Molecule": "molecular",
    "DLPOLY": "trajectory", "PDB": "trajectory", "XYZ": "trajectory", "make_supercell": "trajectory",
    "Input": "io_tools", "Output": "io_tools",
    "compare_properties_dict": "utilities",
    "periodic_table": "tables",
}
_MODULES = ("utilities", "molecular", "trajectory", "io_tools", "tables", "engine", "periodic", "synthetic")
__all__ = sorted(_LAZY)


def __getattr__(name):
    import importlib
    if name in _LAZY:
        return getattr(importlib.import_module("." + _LAZY[name], __name__), name)
    if name in _MODULES:
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
