""":class:`MolecularSystem` and :class:`Molecule` with the reference's public surface
(``pywindow/_internal/molecular.py``: ``Molecule`` :60-551, ``MolecularSystem`` :554-955), computing on
the GPU.

``Molecule.full_analysis()`` issues ONE batched GPU analysis (all stages) instead of the reference's
nine separate calls (which evaluate ``opt_pore_diameter`` three times, molecular.py:192-200); the
individual ``calculate_*`` methods remain available and use the single-purpose entry points.
"""
from __future__ import annotations

import pathlib
from copy import deepcopy

import numpy as np

from . import utilities as U
from .engine import default_engine, record_to_properties
from .io_tools import Input, Output


class _MolecularSystemError(Exception):
    def __init__(self, message):
        self.message = message


class _NotAModularSystemError(Exception):
    def __init__(self, message):
        self.message = message


class Molecule:
    """A single molecule: ``elements``, ``coordinates`` and the growing ``properties`` dict."""

    def __init__(self, mol, system_name, mol_id):
        self._Output = Output()
        self.mol = mol
        self.no_of_atoms = len(mol["elements"])
        self.elements = mol["elements"]
        if "atom_ids" in mol:
            self.atom_ids = mol["atom_ids"]
        self.coordinates = mol["coordinates"]
        self.parent_system = system_name
        self.molecule_id = mol_id
        self.properties = {"no_of_atoms": self.no_of_atoms}
        self._windows = None

    @classmethod
    def load_rdkit_mol(cls, mol, system_name="rdkit", mol_id=0):
        return cls(Input().load_rdkit_mol(mol), system_name, mol_id)

    # ------------------------------------------------------------------ whole analysis, one launch set
    def full_analysis(self, ncpus=1, **kwargs):  # noqa: ARG002
        """molecular.py:156-202.  ``ncpus`` is accepted for compatibility and ignored."""
        X = np.ascontiguousarray(self.coordinates, dtype=np.float64)[None]
        rec = default_engine().analyse_frames(X, self.elements, **kwargs).host_records()[0]
        self._absorb(record_to_properties(rec))
        return self.properties

    def _absorb(self, props):
        self.MW = U.molecular_weight(self.elements)
        self.properties.update(props)
        self.centre_of_mass = props["centre_of_mass"]
        md = props["maximum_diameter"]
        self.maxd_atom_1, self.maxd_atom_2, self.maximum_diameter = md["atom_1"], md["atom_2"], md["diameter"]
        self.average_diameter = props["average_diameter"]
        self.pore_diameter, self.pore_closest_atom = props["pore_diameter"]["diameter"], props["pore_diameter"]["atom"]
        self.pore_volume = props["pore_volume"]
        po = props["pore_diameter_opt"]
        self.pore_diameter_opt, self.pore_opt_closest_atom, self.pore_opt_COM = (
            po["diameter"], po["atom_1"], po["centre_of_mass"])
        self.pore_volume_opt = props["pore_volume_opt"]

    # ------------------------------------------------------------------ single properties
    def calculate_centre_of_mass(self):
        self.centre_of_mass = U.center_of_mass(self.elements, self.coordinates)
        self.properties["centre_of_mass"] = self.centre_of_mass
        return self.centre_of_mass

    def calculate_maximum_diameter(self):
        self.maxd_atom_1, self.maxd_atom_2, self.maximum_diameter = U.max_dim(self.elements, self.coordinates)
        self.properties["maximum_diameter"] = {"diameter": self.maximum_diameter, "atom_1": int(self.maxd_atom_1),
                                               "atom_2": int(self.maxd_atom_2)}
        return self.maximum_diameter

    def calculate_average_diameter(self, **kwargs):
        self.average_diameter = U.find_average_diameter(self.elements, self.coordinates, **kwargs)
        self.properties["average_diameter"] = self.average_diameter
        return self.average_diameter

    def calculate_pore_diameter(self):
        self.pore_diameter, self.pore_closest_atom = U.pore_diameter(self.elements, self.coordinates)
        self.properties["pore_diameter"] = {"diameter": self.pore_diameter, "atom": int(self.pore_closest_atom)}
        return self.pore_diameter

    def calculate_pore_volume(self):
        self.pore_volume = U.sphere_volume(self.calculate_pore_diameter() / 2)
        self.properties["pore_volume"] = self.pore_volume
        return self.pore_volume

    def calculate_pore_diameter_opt(self, **kwargs):
        self.pore_diameter_opt, self.pore_opt_closest_atom, self.pore_opt_COM = U.opt_pore_diameter(
            self.elements, self.coordinates, **kwargs)
        self.properties["pore_diameter_opt"] = {"diameter": self.pore_diameter_opt,
                                                "atom_1": int(self.pore_opt_closest_atom),
                                                "centre_of_mass": self.pore_opt_COM}
        return self.pore_diameter_opt

    def calculate_pore_volume_opt(self, **kwargs):
        self.pore_volume_opt = U.sphere_volume(self.calculate_pore_diameter_opt(**kwargs) / 2)
        self.properties["pore_volume_opt"] = self.pore_volume_opt
        return self.pore_volume_opt

    def calculate_windows(self, ncpus=1, **kwargs):  # noqa: ARG002
        """molecular.py:323-352.  Extra kwargs (``adjust``, ``increment``, ``pore_opt``,
        ``pore_centre``) are passed to :func:`utilities.find_windows`."""
        windows = U.find_windows(self.elements, self.coordinates, **kwargs)
        if windows is not None:
            self.properties.update({"windows": {"diameters": windows[0], "centre_of_mass": windows[1]}})
            return windows[0]
        self.properties.update({"windows": {"diameters": None, "centre_of_mass": None}})
        return None

    def shift_to_origin(self):
        self.coordinates = U.shift_com(self.elements, self.coordinates)
        self._update()

    def molecular_weight(self):
        self.MW = U.molecular_weight(self.elements)
        return float(self.MW)

    # ------------------------------------------------------------------ output
    def dump_properties_json(self, filepath=None, molecular=False, override=False, **kwargs):  # noqa: FBT002
        dict_obj = deepcopy(self.properties)
        if molecular is True:
            dict_obj.update(self.mol)
        if filepath is None:
            filepath = pathlib.Path.cwd() / f"{self.parent_system}_{self.molecule_id}"
        self._Output.dump2json(dict_obj, pathlib.Path(filepath), default=U.to_list, override=override, **kwargs)

    def dump_molecule(self, filepath=None, include_coms=False, override=False, **kwargs):  # noqa: FBT002
        if filepath is None:
            filepath = pathlib.Path.cwd() / f"{self.parent_system}_{self.molecule_id}.pdb"
        filepath = pathlib.Path(filepath)
        key = "elements" if "atom_ids" not in self.mol else "atom_ids"
        mol = self.mol
        if include_coms:
            mol = deepcopy(self.mol)
            extra = [("He", self.properties["centre_of_mass"]),
                     ("Ne", self.properties["pore_diameter_opt"]["centre_of_mass"])]
            wc = self.properties["windows"]["centre_of_mass"]
            if wc is not None:
                extra += [("Ar", c) for c in wc]
            n_ar = 0
            for tag, xyz in extra:
                mol["elements"] = np.concatenate((mol["elements"], np.array([tag])))
                if "atom_ids" in mol:
                    if tag == "Ar":
                        n_ar += 1
                    mol["atom_ids"] = np.concatenate((mol["atom_ids"], np.array([f"Ar{n_ar}" if tag == "Ar" else tag])))
                mol["coordinates"] = np.concatenate((mol["coordinates"], np.array([xyz])))
        self._Output.dump2file(mol, filepath, atom_ids_key=key, override=override, **kwargs)

    def _update(self):
        self.mol["coordinates"] = self.coordinates
        self.calculate_centre_of_mass()
        self.calculate_pore_diameter_opt()


class MolecularSystem:
    """Container for a (possibly periodic, multi-molecule) system; create with ``load_file``,
    ``load_rdkit_mol`` or ``load_system``."""

    def __init__(self):
        self._Input = Input()
        self._Output = Output()
        self.system_id = 0
        self.system = {}
        self.molecules = {}

    @classmethod
    def load_file(cls, filepath):
        obj = cls()
        filepath = pathlib.Path(filepath)
        obj.system = obj._Input.load_file(filepath)
        obj.filename = filepath.name
        obj.system_id = obj.filename.split(".")[0]
        obj.name, _ext = obj.filename.split(".")
        return obj

    @classmethod
    def load_rdkit_mol(cls, mol):
        obj = cls()
        obj.system = obj._Input.load_rdkit_mol(mol)
        return obj

    @classmethod
    def load_system(cls, dict_, system_id="system"):
        obj = cls()
        obj.system = dict_
        obj.system_id = system_id
        return obj

    def rebuild_system(self, override=False):  # noqa: FBT002
        """molecular.py:672-708: supercell + discrete molecules rebuilt across the periodic boundary,
        concatenated back into one system."""
        from .periodic import create_supercell, discrete_molecules
        discrete = discrete_molecules(self.system, rebuild=create_supercell(self.system))
        rebuilt = {
            "coordinates": np.concatenate([m["coordinates"] for m in discrete], axis=0) if discrete
            else np.zeros((0, 3)),
            "atom_ids": np.concatenate([m["atom_ids"] for m in discrete], axis=0) if discrete else np.array([]),
            "elements": np.concatenate([m["elements"] for m in discrete], axis=0) if discrete else np.array([]),
        }
        if override is True:
            self.system.update(rebuilt)
        return self.load_system(rebuilt)

    def swap_atom_keys(self, swap_dict, dict_key="atom_ids"):
        if "atom_ids" not in self.system:
            dict_key = "elements"
        arr = self.system[dict_key]
        for k in range(len(arr)):
            for old, new in swap_dict.items():
                if arr[k] == old:
                    arr[k] = new

    def decipher_atom_keys(self, forcefield="DLF", dict_key="atom_ids"):
        if "atom_ids" not in self.system:
            dict_key = "elements"
        temp = deepcopy(self.system[dict_key])
        cache = {}
        for k in range(len(temp)):
            key = str(temp[k])
            if key not in cache:
                cache[key] = f"{U.decipher_atom_key(key, forcefield=forcefield)}"
            temp[k] = cache[key]
        self.system["elements"] = temp

    def make_modular(self, rebuild=False):  # noqa: FBT002
        """molecular.py:798-824."""
        from .periodic import create_supercell, discrete_molecules
        supercell = create_supercell(self.system) if rebuild is True else None
        dis = discrete_molecules(self.system, rebuild=supercell)
        self.no_of_discrete_molecules = len(dis)
        self.molecules = {i: Molecule(mol=m, system_name=str(self.system_id), mol_id=i) for i, m in enumerate(dis)}

    def system_to_molecule(self):
        return Molecule(mol=self.system, system_name=str(self.system_id), mol_id=0)

    def _modular_system(self):
        if not self.molecules:
            raise _NotAModularSystemError("This system is not modular: run make_modular() first.")
        out = {"elements": np.concatenate([m.mol["elements"] for m in self.molecules.values()]),
               "coordinates": np.concatenate([m.mol["coordinates"] for m in self.molecules.values()])}
        if all("atom_ids" in m.mol for m in self.molecules.values()):
            out["atom_ids"] = np.concatenate([m.mol["atom_ids"] for m in self.molecules.values()])
        return out

    def dump_system(self, filepath=None, modular=False, override=False, **kwargs):  # noqa: FBT002
        if filepath is None:
            filepath = pathlib.Path.cwd() / f"{self.system_id}.pdb"
        system = self._modular_system() if modular else self.system
        key = "atom_ids" if "atom_ids" in system else "elements"
        self._Output.dump2file(system, pathlib.Path(filepath), atom_ids_key=key, override=override, **kwargs)

    def dump_system_json(self, filepath=None, modular=False, override=False, **kwargs):  # noqa: FBT002
        if filepath is None:
            filepath = pathlib.Path.cwd() / f"{self.system_id}"
        system = self._modular_system() if modular else self.system
        self._Output.dump2json(dict(system), pathlib.Path(filepath), default=U.to_list, override=override, **kwargs)
