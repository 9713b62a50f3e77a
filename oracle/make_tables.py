"""TEST/BUILD INFRASTRUCTURE -- export the reference's per-element constants as data files.

The numbers are facts (CCDC elemental radii, DL_FIELD OPLS keys; reference tables.py:22-640) that
the drop-in must share with the reference to give identical results; they are stored as CSV/JSON
data in ``pywindow_b200/data`` rather than as code.  Run in the build container only.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from refharness import load_reference  # noqa: E402

load_reference()
from pywindow._internal import tables as T  # noqa: E402

out = os.path.join(os.path.dirname(HERE), "pywindow_b200", "data")
os.makedirs(out, exist_ok=True)
keys = sorted(set(T.atomic_mass) | set(T.atomic_vdw_radius) | set(T.atomic_covalent_radius),
              key=lambda k: (T.periodic_table.get(k.capitalize(), 999), k))
with open(os.path.join(out, "elements.csv"), "w") as f:
    f.write("symbol,z,mass,vdw_radius,covalent_radius\n")
    for k in keys:
        f.write("{},{},{},{},{}\n".format(
            k, T.periodic_table.get(k.capitalize(), 0),
            repr(T.atomic_mass.get(k, "")), repr(T.atomic_vdw_radius.get(k, "")),
            repr(T.atomic_covalent_radius.get(k, ""))).replace("''", ""))
with open(os.path.join(out, "opls_atom_keys.json"), "w") as f:
    json.dump(T.opls_atom_keys, f, sort_keys=True, separators=(",", ":"))
with open(os.path.join(out, "periodic_table.json"), "w") as f:
    json.dump(T.periodic_table, f, separators=(",", ":"))
print(len(keys), "elements")
