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
    """Device-side result of :func:`discrete_molecules_batch` (views of the engine's pooled buffers: valid until
    the next call with the same batch shape)."""

    def __init__(self, n_mol, mol_off, out_src, out_xyz, status, N, bufs=None):
        self.n_mol_dev, self.mol_off, self.out_src, self.out_xyz, self.status, self.N = n_mol, mol_off, out_src, out_xyz, status, N
        self._bufs = bufs
        self._n_mol = None

    @property
    def n_mol(self):
        if self._n_mol is None:
            self._n_mol = self.n_mol_dev.cpu().numpy()
        return self._n_mol

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


def _dm_buffers(eng, F, N, out_cap, elements):
    """Pooled device buffers of the periodic stage for (F, N, out_cap): nothing is allocated or zero-filled per
    batch (every word that must start at zero is initialised by the kernels themselves)."""
    key = ("dm", F, N, out_cap, tuple(str(e) for e in elements))
    pool = eng._pinned
    b = pool.get(key)
    if b is not None:
        return b
    for k in [k for k in pool if isinstance(k, tuple) and k and k[0] == "dm"][2:]:
        pool.pop(k)                                              # keep at most three shapes alive
    dev, lib = eng.device, eng.lib
    types, cov, mass, term = _type_tables(elements)
    a_types, a_vdw, a_mass = _engine_types(elements)
    mol_cap = N
    e = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)      # noqa: E731
    nbytes = lib.pwb_dm_workspace_bytes(F, N, out_cap)
    b = {"mol_cap": mol_cap, "n_types": len(cov),
         "max_cov": float(max(atomic_covalent_radius[str(x).upper()] for x in set(elements))),
         "types": eng._dev(types, torch.int32), "cov": eng._dev(cov, torch.float64), "mass": eng._dev(mass, torch.float64),
         "term": eng._dev(term, torch.int32),
         "a_types": eng._dev(a_types, torch.int32), "a_vdw": eng._dev(a_vdw, torch.float64), "a_mass": eng._dev(a_mass, torch.float64),
         "n_mol": e(F, torch.int32), "mol_off": e((F, mol_cap + 1), torch.int32), "out_src": e((F, out_cap), torch.int32),
         "out_xyz": e((F, out_cap, 3), torch.float64), "status": e(F, torch.int32),
         "ws": e(nbytes, torch.uint8), "ws_bytes": nbytes,
         # dense batch for the analysis kernels (worst case: every emitted atom its own molecule)
         "item_off": e(F * out_cap + 1, torch.int32), "coords": e((F * out_cap, 3), torch.float64),
         "dense_types": e(F * out_cap, torch.int32), "item_frame": e(F * out_cap, torch.int32),
         "item_mol": e(F * out_cap, torch.int32), "scratch": e(2 * F + 4, torch.int32),
         "counts": torch.zeros(4, dtype=torch.int32).pin_memory(), "d_coords": None}
    pool[key] = b
    return b


def _run_dm(eng, b, d_coords, d_lat, shared, F, N, out_cap, rebuild, tol):
    stream = ctypes.c_void_p(torch.cuda.current_stream(eng.device).cuda_stream)
    _lib.check(eng.lib.pwb_discrete_molecules(
        F, N, int(bool(rebuild)), float(tol), b["max_cov"], d_coords.data_ptr(),
        d_lat.data_ptr() if d_lat is not None else None, shared, b["types"].data_ptr(), b["n_types"],
        b["cov"].data_ptr(), b["mass"].data_ptr(), b["term"].data_ptr(), out_cap, b["mol_cap"], b["n_mol"].data_ptr(),
        b["mol_off"].data_ptr(), b["out_src"].data_ptr(), b["out_xyz"].data_ptr(), b["status"].data_ptr(),
        b["ws"].data_ptr(), b["ws_bytes"], stream))
    return stream


def _compact(eng, b, F, N, out_cap, stream):
    """pwb_dm_compact: dense (item_off, coords, per-atom types, keys) + {items, atoms, largest molecule, status}."""
    _lib.check(eng.lib.pwb_dm_compact(
        F, N, out_cap, b["mol_cap"], b["n_mol"].data_ptr(), b["mol_off"].data_ptr(), b["out_src"].data_ptr(),
        b["out_xyz"].data_ptr(), b["status"].data_ptr(), b["a_types"].data_ptr(), b["item_off"].data_ptr(),
        b["coords"].data_ptr(), b["dense_types"].data_ptr(), b["item_frame"].data_ptr(), b["item_mol"].data_ptr(),
        b["scratch"].data_ptr(), b["counts"].data_ptr(), stream))
    return [int(v) for v in b["counts"].tolist()]


def _check_dm_status(st):
    if st & 1:
        raise RuntimeError("discrete_molecules: an atom has more than 12 bonded neighbours")
    if st & 6:
        raise RuntimeError("discrete_molecules: output capacity exceeded")


def _device_inputs(eng, coords, lattice):
    if isinstance(coords, np.ndarray):       # pageable host array: stage through a pooled pinned buffer
        h = eng.pinned(("dm_in",), coords.shape)
        h.numpy()[...] = coords
        coords = h
    d_coords = coords.to(device=eng.device, dtype=torch.float64, non_blocking=True).contiguous() \
        if isinstance(coords, torch.Tensor) else eng._dev(coords, torch.float64)
    d_lat, shared = None, 1
    if lattice is not None:
        if isinstance(lattice, torch.Tensor):
            shared = 1 if lattice.dim() == 2 else 0
            d_lat = eng._dev(lattice.reshape(-1, 9), torch.float64)
        else:
            lat = np.asarray(lattice, dtype=np.float64)
            shared = 1 if lat.ndim == 2 else 0
            d_lat = eng._dev(lat.reshape(-1, 9), torch.float64)
    return d_coords, d_lat, shared


def discrete_molecules_batch(coords, elements, lattice=None, rebuild=False, tol=0.4, out_factor=2.0, engine=None):
    """``coords`` [F,N,3]; ``lattice`` None | [3,3] | [F,3,3].  Returns a :class:`DiscreteBatch`."""
    eng = engine or default_engine()
    F, N = int(coords.shape[0]), int(coords.shape[1])
    with torch.cuda.device(eng.device):
        d_coords, d_lat, shared = _device_inputs(eng, coords, lattice)
        while True:
            out_cap = int(max(64, out_factor * N))
            b = _dm_buffers(eng, F, N, out_cap, elements)
            stream = _run_dm(eng, b, d_coords, d_lat, shared, F, N, out_cap, rebuild, tol)
            counts = _compact(eng, b, F, N, out_cap, stream)
            if counts[3] & 2 and out_factor < 64:
                out_factor *= 2          # a flood fill outgrew the output buffer: retry larger
                continue
            _check_dm_status(counts[3])
            return DiscreteBatch(b["n_mol"], b["mol_off"], b["out_src"], b["out_xyz"], b["status"], N, b)


def analyse_cells(coords, elements, lattice=None, rebuild=False, tol=0.4, engine=None, out_factor=2.0, key_arrays=False,
                  **kw):
    """Periodic pipeline kept on the device: ``discrete_molecules`` for every cell-frame, a native compaction of
    the molecules found into one dense batch (``pwb_dm_compact``), then ``Molecule.full_analysis`` for every molecule
    (MolecularSystem.make_modular + the per-molecule loop of trajectory.py:512-522 for a whole batch of frames).
    Only ``pwb::`` kernels run between the input copy and the record copy; one 16-byte read-back tells the host how
    many molecules there are.

    ``coords`` [F,N,3]; ``lattice`` None | [3,3] | [F,3,3].  Returns ``(records [M, stride] numpy,
    keys [(frame_index, molecule_index)] (or the two index arrays if ``key_arrays``), sizes [M])`` in the reference's
    molecule order."""
    eng = engine or default_engine()
    F, N = int(coords.shape[0]), int(coords.shape[1])
    with torch.cuda.device(eng.device):
        d_coords, d_lat, shared = _device_inputs(eng, coords, lattice)
        while True:
            out_cap = int(max(64, out_factor * N))
            b = _dm_buffers(eng, F, N, out_cap, elements)
            stream = _run_dm(eng, b, d_coords, d_lat, shared, F, N, out_cap, rebuild, tol)
            items, atoms, biggest, st = _compact(eng, b, F, N, out_cap, stream)
            if st & 2 and out_factor < 64:
                out_factor *= 2
                continue
            _check_dm_status(st)
            break
        if items == 0:
            return np.zeros((0, eng.rec_stride)), [], np.zeros(0, dtype=np.int64)
        out = eng.analyse(b["coords"][:atoms], b["item_off"][:items + 1], b["dense_types"][:atoms], b["a_vdw"], b["a_mass"],
                          shared_types=False, max_atoms=biggest, **kw)
        # records through a pooled pinned buffer (a pageable read-back runs at ~2 GB/s), then one host copy
        h = eng.pinned(("dm_rec",), (items, eng.rec_stride))
        h.copy_(out.records, non_blocking=True)
        torch.cuda.current_stream(eng.device).synchronize()
        rec = h.numpy().copy()
        frame_idx = b["item_frame"][:items].cpu().numpy()
        mol_idx = b["item_mol"][:items].cpu().numpy()
        sizes = np.diff(b["item_off"][:items + 1].cpu().numpy())
        if key_arrays:               # (frame index, molecule index) as two int arrays: no per-item Python objects
            return rec, (frame_idx.astype(np.int64), mol_idx.astype(np.int64)), sizes
        keys = list(zip(frame_idx.tolist(), mol_idx.tolist()))      # python ints, without a numpy scalar per entry
        return rec, keys, sizes


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
