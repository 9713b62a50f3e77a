"""Per-element constants used by the pore-geometry path.

Same data as the reference's ``pywindow/_internal/tables.py`` (``atomic_mass`` :22, ``atomic_vdw_radius``
:111, ``atomic_covalent_radius`` :200, ``opls_atom_keys`` :290, ``periodic_table`` :643), kept as data
files under ``pywindow_b200/data`` (exported by ``oracle/make_tables.py``) and exposed under the
reference's names.  Keys of the three radius/mass dicts are upper-case element symbols.
"""
from __future__ import annotations

import csv
import json
import os

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

atomic_mass: dict[str, float] = {}
atomic_vdw_radius: dict[str, float] = {}
atomic_covalent_radius: dict[str, float] = {}

with open(os.path.join(_DATA, "elements.csv")) as _f:
    for _row in csv.DictReader(_f):
        _k = _row["symbol"]
        if _row["mass"]:
            atomic_mass[_k] = float(_row["mass"])
        if _row["vdw_radius"]:
            atomic_vdw_radius[_k] = float(_row["vdw_radius"])
        if _row["covalent_radius"]:
            atomic_covalent_radius[_k] = float(_row["covalent_radius"])

with open(os.path.join(_DATA, "opls_atom_keys.json")) as _f:
    opls_atom_keys: dict[str, list[str]] = json.load(_f)
with open(os.path.join(_DATA, "periodic_table.json")) as _f:
    periodic_table: dict[str, int] = json.load(_f)
