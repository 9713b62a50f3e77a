"""File readers / writers used on either side of the analysis path (XYZ, PDB, MOL V3000, JSON).

Host text IO only; mirrors the behaviour of the reference's ``pywindow/_internal/io_tools.py``
(``Input`` :42-205, ``Output`` :208-493) for the formats its tests and examples use.
"""
from __future__ import annotations

import json
import pathlib

import numpy as np

from .utilities import unit_cell_to_lattice_array


class _CorruptedPDBFileError(Exception):
    def __init__(self, message):
        self.message = message


class _CorruptedXYZFileError(Exception):
    def __init__(self, message):
        self.message = message


class _NotADictionaryError(Exception):
    def __init__(self, message):
        self.message = message


class _FileTypeError(Exception):
    def __init__(self, message):
        self.message = message


def _is_atom(line):
    return line[:6] in ("HETATM", "ATOM  ")


class Input:
    """``Input().load_file(path)`` -> system dict (elements, coordinates, [atom_ids, unit_cell, lattice])."""

    def __init__(self):
        self.system = {}
        self._readers = {".xyz": self._read_xyz, ".pdb": self._read_pdb, ".mol": self._read_mol}

    def load_file(self, filepath):
        self.file_path = pathlib.Path(filepath)
        self.file_type = self.file_path.suffix
        self.file_name = self.file_path.name
        with self.file_path.open("r") as fh:
            self.file_content = fh.readlines()
        if self.file_type not in self._readers:
            raise _FileTypeError(f"The {self.file_type} file extension is not supported (xyz, pdb, mol).")
        return self._readers[self.file_type]()

    def load_rdkit_mol(self, mol):
        """rdkit Mol -> system dict (io_tools.py:80-104); needs rdkit's mol-block writer."""
        from rdkit import Chem  # noqa: PLC0415  (optional dependency, absent in this image)
        self.file_content = Chem.MolToMolBlock(mol, forceV3000=True).split("\n")
        self.file_content = [ln + "\n" for ln in self.file_content]
        return self._read_mol()

    def _read_xyz(self):
        body = self.file_content[2:]
        try:
            self.file_remarks = self.file_content[1]
            tokens = [ln.split() for ln in body]
            self.system = {"elements": np.array([t[0] for t in tokens]),
                           "coordinates": np.array([[float(t[1]), float(t[2]), float(t[3])] for t in tokens])}
        except IndexError:
            raise _CorruptedXYZFileError(
                "The XYZ file is corrupted in some way. For example, an empty line at the end etc. or it is a "
                "trajectory. If the latter is the case, please use `trajectory` module, otherwise fix it.") from None
        return self.system

    def _read_pdb(self):
        lines = self.file_content
        if sum(ln.count("END ") for ln in lines) > 1:
            raise _CorruptedPDBFileError(
                "Multiple 'END' statements were found in this PDB file.If this is a trajectory, use a trajectory "
                "module, Otherwise, fix it.")
        atoms = [ln for ln in lines if _is_atom(ln)]
        cell = [float(ln[a:b]) for ln in lines if ln[:6] == "CRYST1"
                for a, b in ((6, 15), (15, 24), (24, 33), (33, 40), (40, 47), (47, 54))]
        self.system = {"remarks": [ln for ln in lines if ln[:6] == "REMARK"], "unit_cell": np.array(cell)}
        if self.system["unit_cell"].any():
            self.system["lattice"] = unit_cell_to_lattice_array(self.system["unit_cell"])
        self.system["atom_ids"] = np.array([ln[12:16].strip() for ln in atoms], dtype="<U8")
        self.system["elements"] = np.array([ln[76:78].strip() for ln in atoms], dtype="<U8")
        self.system["coordinates"] = np.array([[float(ln[30:38]), float(ln[38:46]), float(ln[46:54])] for ln in atoms])
        return self.system

    def _read_mol(self):
        self.system = {}
        if self.file_content[2] != "\n":
            self.system["remarks"] = self.file_content[2]
        elements, coords, inside = [], [], False
        for tok in (ln.split() for ln in self.file_content):
            if len(tok) > 2:
                if tok[2] == "END" and tok[3] == "ATOM":
                    inside = False
                if inside:
                    elements.append(tok[3])
                    coords.append(tok[4:7])
                if tok[2] == "BEGIN" and tok[3] == "ATOM":
                    inside = True
        self.system["elements"] = np.array(elements)
        self.system["coordinates"] = np.array(coords, dtype=float)
        return self.system


class Output:
    """JSON / XYZ / PDB writers (io_tools.py:208-493)."""

    def __init__(self):
        self.cwd = pathlib.Path.cwd()

    def dump2json(self, obj, filepath, default=None, override=False, **kwargs):  # noqa: FBT002
        if not isinstance(obj, dict):
            raise _NotADictionaryError("This function only accepts dictionaries as input")
        filepath = pathlib.Path(filepath)
        if filepath.suffix != ".json":
            filepath = pathlib.Path(str(filepath) + ".json")
        if not override and filepath.is_file():
            raise FileExistsError(f"The file {filepath} already exists. Use a different filepath, or set "
                                  "the 'override' kwarg to True.")
        with filepath.open("w+") as fh:
            json.dump(obj, fh, default=default, **kwargs)

    def dump2file(self, obj, filepath, atom_ids_key="elements", override=False, **kwargs):  # noqa: FBT002
        """io_tools.py:267-347: XYZ or PDB by extension; kwargs as in the reference (elements_key, coordinates_key,
        remarks, cryst, space_group, forcefield, decipher, resname, chainid, resseq)."""
        filepath = pathlib.Path(filepath)
        if not override and filepath.is_file():
            raise FileExistsError(f"The file {filepath} already exists. Use a different filepath, or set "
                                  "the 'override' kwarg to True.")
        if filepath.suffix == ".pdb":
            self._save_pdb(obj, filepath, atom_ids_key=atom_ids_key, **kwargs)
        elif filepath.suffix == ".xyz":
            for k in ("cryst", "space_group", "resname", "chainid", "resseq"):
                kwargs.pop(k, None)
            self._save_xyz(obj, filepath, **kwargs)
        else:
            raise _FileTypeError(f"The {filepath.suffix} file extension is not supported for dumping a MolecularSystem "
                                 "or a Molecule. Please use XYZ or PDB.")

    @staticmethod
    def _elements(system, key, forcefield, decipher):
        el = system[key]
        if decipher is True:
            if forcefield is None:
                raise ValueError("forcefield must be provided when decipher is True")
            from .utilities import decipher_atom_key
            el = np.array([decipher_atom_key(k, forcefield=forcefield) for k in el])
        return el

    def _save_xyz(self, system, filepath, elements_key="elements", coordinates_key="coordinates", remarks=None,
                  forcefield=None, decipher=False):  # noqa: FBT002
        """Byte-for-byte the file of io_tools.py:349-383: count, remark line, ``El x y z`` with two decimals."""
        if isinstance(remarks, (list, tuple)):
            remarks = ";".join(remarks)
        elif remarks is None:
            remarks = ""
        el = self._elements(system, elements_key, forcefield, decipher)
        body = "".join("{} {:.2f} {:.2f} {:.2f}\n".format(e, *c) for e, c in zip(el, system[coordinates_key]))
        pathlib.Path(filepath).write_text(f"{len(el):0d}\n{remarks!s}\n" + body)

    def _save_pdb(self, system, filepath, atom_ids_key="elements", elements_key="elements", coordinates_key="coordinates",
                  remarks=None, cryst="unit_cell", space_group=None, forcefield=None, decipher=False, resname="MOL",  # noqa: FBT002
                  chainid="A", resseq=1):
        """Byte-for-byte the file of io_tools.py:385-493: generator remark, optional remarks, CRYST1 (P1 unless a
        space group is given), one ATOM record per atom (serial, centred 4-character name, residue, chain, sequence
        number, 8.3f coordinates, right-justified element), END without a trailing newline."""
        filepath = pathlib.Path(filepath)
        lines = ["REMARK File generated using pyWINDOW."]
        if isinstance(remarks, (list, tuple)):
            lines += [f"REMARK {r}" for r in remarks]
        elif isinstance(remarks, (str, int, float)):
            lines.append(f"REMARK {remarks}")
        if cryst in system and np.asarray(system[cryst]).any():
            c = system[cryst]
            lines.append("CRYST1" + "".join(f"{v:9.3f}" for v in c[:3]) + "".join(f"{v:7.2f}" for v in c[3:])
                         + " " + ("P1" if space_group is None else space_group))
        ids = system[atom_ids_key]
        el = self._elements(system, elements_key, forcefield, decipher)
        xyz = system[coordinates_key]
        seq = str(resseq).rjust(4)
        for i in range(len(list(ids))):
            lines.append(f"ATOM  {i + 1:5d} {ids[i].center(4):4} {resname:3} {chainid}{seq}    "
                         f"{xyz[i][0]:8.3f}{xyz[i][1]:8.3f}{xyz[i][2]:8.3f}{' ' * 22}{el[i].rjust(2):2}  ")
        lines.append("END")
        if filepath.suffix != ".pdb":
            filepath = pathlib.Path(f"{filepath}.pdb")
        filepath.write_text("\n".join(lines))
