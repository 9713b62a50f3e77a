"""Periodic cells: ``create_supercell`` and ``discrete_molecules`` (reference utilities.py:768-810,
:820-1085) on the GPU.

The GPU never materialises the 27-image supercell: image coordinates are generated on the fly from
fractional coordinates inside the bond-table and flood-fill kernels (``csrc/pwb_periodic.cu``).
``create_supercell`` still returns the reference's dictionary for API compatibility (vectorised numpy,
host side, not on the per-frame path).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from . import utilities as U
from .engine import default_engine
from .tables import atomic_covalent_radius, atomic_mass

TERMINAL = ("H", "CL", "BR", "F", "HE", "AR", "NE", "KR", "XE", "RN")   # utilities.py:933


def create_supercell(system, supercell=None):
    """utilities.py:768-810 (host, vectorised): 27 (or user-defined) images, a outermost."""
    if supercell is None:
        supercell = [[-1, 1], [-1, 1], [-1, 1]]
    matrix = U.unit_cell_to_lattice_array(system["unit_cell"]) if "lattice" not in system else system["lattice"]
    X = np.asarray(system["coordinates"], dtype=np.float64)
    frac = U.cart2frac_all(X, matrix)
    shifts = [(a, b, c) for a in range(supercell[0][0], supercell[0][1] + 1)
              for b in range(supercell[1][0], supercell[1][1] + 1)
              for c in range(supercell[2][0], supercell[2][1] + 1)]
    coords = np.concatenate([U.frac2cart_all(frac + np.array(s, dtype=np.float64), matrix) for s in shifts], axis=0)
    ids = system["atom_ids"] if "atom_ids" in system else system["elements"]
    return {"elements": np.concatenate([system["elements"]] * len(shifts)),
            "atom_ids": np.concatenate([ids] * len(shifts)),
            "coordinates": coords,
            "unit_cell": U.lattice_array_to_unit_cell(matrix),
            "lattice": matrix}


def _type_tables(elements):
    up = [str(e).upper() for e in elements]
    uniq = sorted(set(up))
    idx = {e: i for i, e in enumerate(uniq)}
    types = np.array([idx[e] for e in up], dtype=np.int32)
    cov = np.array([atomic_covalent_radius[e] for e in uniq], dtype=np.float64)
    mass = np.array([atomic_mass[e] for e in uniq], dtype=np.float64)
    term = np.array([e in TERMINAL for e in uniq], dtype=np.int32)
    return types, cov, mass, term


class DiscreteBatch:
    """Device-side result of :func:`discrete_molecules_batch`."""

    def __init__(self, n_mol, mol_off, out_src, out_xyz, status, N):
        self.n_mol, self.mol_off, self.out_src, self.out_xyz, self.status, self.N = n_mol, mol_off, out_src, out_xyz, status, N

    def molecules(self, frame, elements, atom_ids=None):
        """Frame's molecules as the reference's list of dicts (utilities.py:1046-1085)."""
        nm = int(self.n_mol[frame])
        off = self.mol_off[frame].cpu().numpy()
        src = self.out_src[frame].cpu().numpy()
        xyz = self.out_xyz[frame].cpu().numpy()
        out = []
        for k in range(nm):
            s = src[off[k]:off[k + 1]]
            d = {"elements": np.array(np.asarray(elements)[s], dtype="str"), "coordinates": xyz[off[k]:off[k + 1]].copy()}
            if atom_ids is not None:
                d["atom_ids"] = np.array(np.asarray(atom_ids)[s], dtype="str")
            out.append(d)
        return out


def discrete_molecules_batch(coords, elements, lattice=None, rebuild=False, tol=0.4, out_factor=2.0, engine=None):
    """``coords`` [F,N,3]; ``lattice`` None | [3,3] | [F,3,3].  Returns a :class:`DiscreteBatch`."""
    eng = engine or default_engine()
    lib, dev = eng.lib, eng.device
    F, N = int(coords.shape[0]), int(coords.shape[1])
    types, cov, mass, term = _type_tables(elements)
    max_cov = float(max(atomic_covalent_radius[str(e).upper()] for e in set(elements)))
    with torch.cuda.device(dev):
        d_coords = eng._dev(coords, torch.float64)
        d_lat, shared = None, 1
        if lattice is not None:
            lat = np.asarray(lattice, dtype=np.float64)
            shared = 1 if lat.ndim == 2 else 0
            d_lat = eng._dev(lat.reshape(-1, 9), torch.float64)
        d_types, d_cov, d_mass, d_term = (eng._dev(types, torch.int32), eng._dev(cov, torch.float64),
                                          eng._dev(mass, torch.float64), eng._dev(term, torch.int32))
        while True:
            out_cap, mol_cap = int(max(64, out_factor * N)), N
            n_mol = torch.zeros(F, dtype=torch.int32, device=dev)
            mol_off = torch.zeros((F, mol_cap + 1), dtype=torch.int32, device=dev)
            out_src = torch.zeros((F, out_cap), dtype=torch.int32, device=dev)
            out_xyz = torch.zeros((F, out_cap, 3), dtype=torch.float64, device=dev)
            status = torch.zeros(F, dtype=torch.int32, device=dev)
            nbytes = lib.pwb_dm_workspace_bytes(F, N, out_cap)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(lib.pwb_discrete_molecules(
                F, N, int(bool(rebuild)), float(tol), max_cov, d_coords.data_ptr(),
                d_lat.data_ptr() if d_lat is not None else None, shared, d_types.data_ptr(), len(cov),
                d_cov.data_ptr(), d_mass.data_ptr(), d_term.data_ptr(), out_cap, mol_cap, n_mol.data_ptr(),
                mol_off.data_ptr(), out_src.data_ptr(), out_xyz.data_ptr(), status.data_ptr(), ws.data_ptr(),
                nbytes, stream))
            st = int(status.max().item())
            if st & 2 and out_factor < 64:
                out_factor *= 2          # a flood fill outgrew the output buffer: retry larger
                continue
            if st & 1:
                raise RuntimeError("discrete_molecules: an atom has more than 12 bonded neighbours")
            if st & 6:
                raise RuntimeError("discrete_molecules: output capacity exceeded")
            return DiscreteBatch(n_mol.cpu().numpy(), mol_off, out_src, out_xyz, status, N)


def analyse_cells(coords, elements, lattice=None, rebuild=False, tol=0.4, engine=None, **kw):
    """Periodic pipeline kept on the device: ``discrete_molecules`` for every cell-frame, then
    ``Molecule.full_analysis`` for every molecule found (MolecularSystem.make_modular + the per-molecule
    loop of trajectory.py:512-522 for a whole batch of frames).

    ``coords`` [F,N,3]; ``lattice`` None | [3,3] | [F,3,3].  Returns ``(records [M, stride] numpy,
    keys [(frame_index, molecule_index)], sizes [M])`` in the reference's molecule order."""
    eng = engine or default_engine()
    dev = eng.device
    res = discrete_molecules_batch(coords, elements, lattice=lattice, rebuild=rebuild, tol=tol, engine=eng)
    F = int(coords.shape[0])
    with torch.cuda.device(dev):
        nm = torch.as_tensor(res.n_mol, device=dev, dtype=torch.int64)
        mol_cap = res.mol_off.shape[1] - 1
        out_cap = res.out_src.shape[1]
        mmask = torch.arange(mol_cap, device=dev)[None, :] < nm[:, None]
        starts = res.mol_off[:, :-1][mmask].to(torch.int64)
        ends = res.mol_off[:, 1:][mmask].to(torch.int64)
        sizes = ends - starts
        if sizes.numel() == 0:
            return np.zeros((0, eng.rec_stride)), [], np.zeros(0, dtype=np.int64)
        tot = torch.gather(res.mol_off.to(torch.int64), 1, nm[:, None]).squeeze(1)
        amask = torch.arange(out_cap, device=dev)[None, :] < tot[:, None]
        dense_xyz = res.out_xyz[amask]
        types, vdw, mass = _engine_types(elements)
        dense_types = torch.as_tensor(types, device=dev)[res.out_src[amask].to(torch.int64)]
        item_off = torch.zeros(sizes.numel() + 1, dtype=torch.int64, device=dev)
        item_off[1:] = torch.cumsum(sizes, 0)
        max_atoms = int(sizes.max().item())
        out = eng.analyse(dense_xyz, item_off.to(torch.int32), dense_types, vdw, mass, shared_types=False,
                          max_atoms=max_atoms, **kw)
        frame_idx = torch.arange(F, device=dev)[:, None].expand(F, mol_cap)[mmask].cpu().numpy()
        mol_idx = torch.arange(mol_cap, device=dev)[None, :].expand(F, mol_cap)[mmask].cpu().numpy()
        keys = list(zip(frame_idx.tolist(), mol_idx.tolist()))      # python ints, without a numpy scalar per entry
        return out.host_records(), keys, sizes.cpu().numpy()


def _engine_types(elements):
    from .engine import element_types
    return element_types(elements)


def discrete_molecules(system, rebuild=None, tol=0.4):
    """utilities.py:820-1085 for one system dict -> list of molecule dicts.  ``rebuild`` is the
    supercell dict of :func:`create_supercell` in the reference; only its presence matters here."""
    try:
        elements = system["elements"]
    except KeyError:
        raise U._FunctionError(
            "The 'elements' key is missing in the 'system' dictionary attribute of the MolecularSystem object. "
            "Which means, you need to decipher the forcefield based atom keys first (see manual).") from None
    X = np.ascontiguousarray(system["coordinates"], dtype=np.float64)
    lattice = None
    if "lattice" in system and np.shape(system["lattice"]) == (3, 3):
        lattice = np.asarray(system["lattice"], dtype=np.float64)
    elif "unit_cell" in system and np.shape(system["unit_cell"]) == (6,):
        lattice = U.unit_cell_to_lattice_array(system["unit_cell"])
    if rebuild is not None and lattice is None:
        lattice = np.asarray(rebuild["lattice"], dtype=np.float64)
    res = discrete_molecules_batch(X[None], elements, lattice=lattice, rebuild=rebuild is not None, tol=tol)
    return res.molecules(0, elements, system.get("atom_ids"))
