"""Trajectory front-ends (DL_POLY HISTORY, XYZ, PDB) with the reference's interface
(``pywindow/_internal/trajectory.py``: ``Trajectory.analysis`` :350-494, ``DLPOLY`` :589-833, ``XYZ``
:836-931, ``PDB`` :934-1045) and a *batched* GPU analysis.

Where the reference loops ``frame -> MolecularSystem -> Molecule.full_analysis`` (optionally fanned out
over a ``multiprocessing.Pool``), ``analysis()`` here decodes the requested frames into one
``[F, N, 3]`` array, analyses them in large GPU batches and rebuilds the same
``analysis_output[frame][molecule] = properties`` dictionaries.  If ``torch.distributed`` is
initialised, the frame list is sharded over the ranks (contiguous blocks) and the result records are
gathered once at the end (NCCL over NVLink on GPUs).

Known reference bugs at this boundary are superseded, not reproduced: ``frames=(a, b)`` works,
PDB trajectory frames decode, XYZ frames are delimited by their atom-count lines.
"""
from __future__ import annotations

import json
import gc
import mmap
import pathlib
import re

import numpy as np

from . import _lib
from . import utilities as U
from .distributed import gather_records, shard_range, world
from .engine import default_engine, record_to_properties
from .molecular import MolecularSystem


def _host_threads():
    """Decoder threads for this process: all cores, shared evenly between the ranks of one node."""
    import os
    local = int(os.environ.get("LOCAL_WORLD_SIZE", "0") or 0)
    if local <= 0:
        _, local = world()
    return max(1, (os.cpu_count() or 1) // max(1, local))


class _ParallelAnalysisError(Exception):
    def __init__(self, message):
        self.message = message


class _TrajectoryError(Exception):
    def __init__(self, message):
        self.message = message


class _FormatError(Exception):
    def __init__(self, message):
        self.message = message


class _FunctionError(Exception):
    def __init__(self, message):
        self.message = message


def make_supercell(system, supercell=None):
    """Supercell of a periodic ``system`` dict as a new :class:`MolecularSystem` (trajectory.py:75-101):
    images 1..s along a, b and c.  (The reference uses ``supercell[1]`` for the c axis as well; the c
    entry is honoured here.)"""
    from .periodic import create_supercell
    if supercell is None:
        supercell = [1, 1, 1]
    user = [[1, int(supercell[0])], [1, int(supercell[1])], [1, int(supercell[2])]]
    return MolecularSystem.load_system(create_supercell(system=system, supercell=user))


class LazyFrame(dict):
    """``{molecule_key: properties}`` of one analysed frame, built from the GPU result records on first
    access.  ``analysis_output`` stays a plain ``dict`` of these, so a 100 000-frame analysis does not
    pay for 100 000 Python property dictionaries nobody looks at (trajectory.py:483, :516-522 build
    them eagerly)."""

    __slots__ = ("_pending",)

    def __init__(self, pending=()):
        super().__init__()
        self._pending = list(pending)

    def _add(self, mol_key, records, row):
        # (key, record table, row): the row view is only taken when somebody looks (slicing 100 000 numpy rows
        # costs more than analysing them)
        self._pending.append((mol_key, records, row))

    def _load(self):
        if self._pending:
            pend, self._pending = self._pending, []
            for k, records, row in pend:
                super().__setitem__(k, record_to_properties(records[row]))
        return self

    def __getitem__(self, k):
        return dict.__getitem__(self._load(), k)

    def __iter__(self):
        return dict.__iter__(self._load())

    def __len__(self):
        return dict.__len__(self) + len({p[0] for p in self._pending if not dict.__contains__(self, p[0])})

    def __contains__(self, k):
        return dict.__contains__(self, k) or any(p[0] == k for p in self._pending)

    def __eq__(self, other):
        return dict.__eq__(self._load(), other)

    def __ne__(self, other):
        return not self.__eq__(other)

    def __repr__(self):
        return dict.__repr__(self._load())

    def keys(self):
        return dict.keys(self._load())

    def values(self):
        return dict.values(self._load())

    def items(self):
        return dict.items(self._load())

    def get(self, k, default=None):
        return dict.get(self._load(), k, default)

    def copy(self):
        return dict(self._load())


class FrameMap(dict):
    """``Trajectory.trajectory_map``: ``{frame: [start_byte, end_byte]}`` (plus ``"header"`` for HISTORY files,
    trajectory.py:647-689) backed by one int64 ``[F, 2]`` array.  The dictionary entries are only created when
    somebody looks at the dictionary; the analysis path reads the array (``bounds``) -- 1e5 two-element lists cost
    as much as decoding 3e4 frames."""

    __slots__ = ("pairs", "_loaded")

    def __init__(self, pairs=None, **extra):
        super().__init__(**extra)
        self.pairs = np.zeros((0, 2), dtype=np.int64) if pairs is None else np.ascontiguousarray(pairs, dtype=np.int64)
        self._loaded = self.pairs.shape[0] == 0

    def bounds(self, frames):
        """(starts, ends) int64 arrays of the given frames."""
        idx = np.asarray(frames, dtype=np.int64)
        if idx.size and (idx.min() < 0 or idx.max() >= self.pairs.shape[0]):
            raise KeyError(int(idx.max() if idx.max() >= self.pairs.shape[0] else idx.min()))
        sel = self.pairs[idx]
        return np.ascontiguousarray(sel[:, 0]), np.ascontiguousarray(sel[:, 1])

    def _load(self):
        if not self._loaded:
            self._loaded = True
            dict.update(self, enumerate(self.pairs.tolist()))
        return self

    def __missing__(self, k):                       # single look-ups do not materialise the whole map
        if isinstance(k, (int, np.integer)) and 0 <= k < self.pairs.shape[0]:
            return self.pairs[k].tolist()
        raise KeyError(k)

    def __iter__(self):
        return dict.__iter__(self._load())

    def __len__(self):
        return dict.__len__(self._load())

    def __contains__(self, k):
        return dict.__contains__(self._load(), k)

    def __repr__(self):
        return dict.__repr__(self._load())

    def __eq__(self, other):
        return dict.__eq__(self._load(), other)

    def __ne__(self, other):
        return not self.__eq__(other)

    def keys(self):
        return dict.keys(self._load())

    def values(self):
        return dict.values(self._load())

    def items(self):
        return dict.items(self._load())

    def get(self, k, default=None):
        return dict.get(self._load(), k, default)

    def copy(self):
        return dict(self._load())


def _bounds(traj, frames):
    """(starts, ends) of ``frames`` from a FrameMap or a plain dict map."""
    tm = traj.trajectory_map
    if isinstance(tm, FrameMap):
        return tm.bounds(frames)
    return (np.array([tm[f][0] for f in frames], dtype=np.int64), np.array([tm[f][1] for f in frames], dtype=np.int64))


class RecordKeys:
    """(frame, molecule key) of every row of a record table as two int64 arrays -- molecule key -1 stands for the
    string key ``"0"`` of a non-modular analysis (trajectory.py:483).  Behaves like the list of tuples it replaces
    (``len``, iteration, indexing, ``==``) without creating 1e5 tuples unless somebody iterates."""

    __slots__ = ("frames", "mols")

    def __init__(self, frames, mols):
        self.frames = np.ascontiguousarray(frames, dtype=np.int64)
        self.mols = np.ascontiguousarray(mols, dtype=np.int64)

    @classmethod
    def single(cls, frames):
        f = np.asarray(frames, dtype=np.int64)
        return cls(f, np.full(f.shape, -1, dtype=np.int64))

    @classmethod
    def concat(cls, parts):
        parts = [p if isinstance(p, RecordKeys) else cls.from_pairs(p) for p in parts]
        if not parts:
            return cls(np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64))
        return cls(np.concatenate([p.frames for p in parts]), np.concatenate([p.mols for p in parts]))

    @classmethod
    def from_pairs(cls, pairs):
        pairs = list(pairs)
        return cls([p[0] for p in pairs], [-1 if isinstance(p[1], str) else p[1] for p in pairs])

    def __len__(self):
        return int(self.frames.shape[0])

    def __iter__(self):
        return iter(zip(self.frames.tolist(), ["0" if m < 0 else m for m in self.mols.tolist()]))

    def __getitem__(self, i):
        if isinstance(i, slice):
            return RecordKeys(self.frames[i], self.mols[i])
        m = int(self.mols[i])
        return (int(self.frames[i]), "0" if m < 0 else m)

    def __eq__(self, other):
        return list(self) == list(other)


class AnalysisOutput(dict):
    """``Trajectory.analysis_output``: ``{frame: {molecule_key: properties}}`` (trajectory.py:483, :516-522).  The
    batched analysis hands over whole record tables; they are parked as blocks and only turned into per-frame
    :class:`LazyFrame` entries when somebody first looks at the dictionary, so ``analysis()`` itself does no
    per-frame Python work (100 000 small containers cost twice the GPU time of analysing the frames)."""

    __slots__ = ("_blocks",)

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self._blocks = []

    def _add_block(self, records, keys):
        self._blocks.append((records, keys))

    def _load(self):
        if self._blocks:
            blocks, self._blocks = self._blocks, []
            # small containers only: the cyclic collector would rescan the growing heap every few hundred
            # allocations (8x the cost of the loop itself), and nothing allocated here can form a cycle
            gc_was_on = gc.isenabled()
            gc.disable()
            try:
                for records, keys in blocks:
                    # every analysed frame REPLACES its entry (analysis_output[frame] = results): with
                    # override=True no stale molecule keys of an earlier run survive
                    fresh = set()
                    for row, (frame, mol_key) in enumerate(keys):
                        if frame not in fresh:
                            fresh.add(frame)
                            dict.__setitem__(self, frame, LazyFrame())
                        dict.__getitem__(self, frame)._add(mol_key, records, row)
            finally:
                if gc_was_on:
                    gc.enable()
        return self

    def __getitem__(self, k):
        return dict.__getitem__(self._load(), k)

    def __setitem__(self, k, v):
        dict.__setitem__(self._load(), k, v)

    def __delitem__(self, k):
        dict.__delitem__(self._load(), k)

    def __iter__(self):
        return dict.__iter__(self._load())

    def __len__(self):
        return dict.__len__(self._load())

    def __contains__(self, k):
        return dict.__contains__(self._load(), k)

    def __eq__(self, other):
        return dict.__eq__(self._load(), other)

    def __ne__(self, other):
        return not self.__eq__(other)

    def __repr__(self):
        return dict.__repr__(self._load())

    def __bool__(self):
        return bool(self._blocks) or dict.__len__(self) > 0

    def keys(self):
        return dict.keys(self._load())

    def values(self):
        return dict.values(self._load())

    def items(self):
        return dict.items(self._load())

    def get(self, k, default=None):
        return dict.get(self._load(), k, default)

    def pop(self, *a):
        return dict.pop(self._load(), *a)

    def copy(self):
        return dict(self._load())

    def record_table(self):
        """(records [M, stride] float64, keys [(frame, molecule_key)]) of everything analysed by the batched path
        and not yet looked at through the dictionary interface -- bulk post-processing without any dicts."""
        if not self._blocks:
            return np.zeros((0, 0)), RecordKeys.concat([])
        return (np.concatenate([b[0] for b in self._blocks]), RecordKeys.concat([b[1] for b in self._blocks]))


class Trajectory:
    """Shared machinery: frame selection, batched GPU analysis, result storage."""

    batch_frames = 16384          # molecule-frames per GPU batch of the single-molecule path
    modular_batch_cells = 2048    # cell-frames per GPU batch of the modular (periodic) path

    def __init__(self, filepath):
        self.filepath = pathlib.Path(filepath)
        self.system_id = self.filepath.name.split(".")[0]
        self.frames = {}
        self.analysis_output = AnalysisOutput()
        self.trajectory_map = {}
        self.no_of_frames = 0

    # ---- the file stays mapped for the lifetime of the object: mapping and (above all) unmapping a multi-GB file
    # once per decoded batch costs as much as decoding the batch
    def _mapped(self):
        """(mmap, uint8 numpy view) of the trajectory file, opened on first use."""
        mm = self.__dict__.get("_mm")
        if mm is None:
            with self.filepath.open("rb") as fh:
                mm = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ)
            self._mm, self._mm_view = mm, np.frombuffer(mm, dtype=np.uint8)
        return self._mm, self._mm_view

    def close(self):
        """Release the file mapping (it is re-opened on demand)."""
        mm = self.__dict__.pop("_mm", None)
        self.__dict__.pop("_mm_view", None)
        if mm is not None:
            try:
                mm.close()
            except BufferError:          # a numpy view handed out earlier is still alive: the GC will unmap
                pass

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- to be provided by the format classes
    def _decode_frame(self, frame_bytes):
        raise NotImplementedError

    def _frame_bytes(self, frame):
        start, end = self.trajectory_map[frame]
        with self.filepath.open("rb") as fh, mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ) as mm:
            return mm[start:end]

    # ---- frame access (trajectory.py:121-249)
    def _get_frame(self, frame_coordinates, frame_no, swap_atoms=None, forcefield=None, extract_data=True):  # noqa: FBT002
        start, end = frame_coordinates
        with self.filepath.open("rb") as fh, mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ) as mm:
            raw = mm[start:end]
        if extract_data is False:
            return raw.decode("utf-8")
        molsys = MolecularSystem.load_system(self._decode_frame(raw), "_".join([self.system_id, str(frame_no)]))
        if swap_atoms is not None:
            molsys.swap_atom_keys(swap_atoms)
        if forcefield is not None:
            molsys.decipher_atom_keys(forcefield)
        return molsys

    def _select(self, frames):
        if isinstance(frames, (int, np.integer)):
            return [int(frames)]
        if isinstance(frames, list):
            if not all(isinstance(f, (int, np.integer)) for f in frames):
                raise _FunctionError("The list should be populated with integers only.")
            return [int(f) for f in frames]
        if isinstance(frames, tuple):
            if len(frames) != 2 or not all(isinstance(f, (int, np.integer)) for f in frames):
                raise _FunctionError("The tuple should contain only two integers for the begining and the end "
                                     "of the frames range.")
            return list(range(frames[0], frames[1]))
        if isinstance(frames, str):
            if frames in ("all", "everything"):
                return list(range(self.no_of_frames))
            raise _FunctionError("Didn't recognise the keyword. (see manual)")
        raise _FunctionError("frames must be an int, a list of ints, a (start, end) tuple or 'all'.")

    def get_frames(self, frames="all", override=False, **kwargs):  # noqa: FBT002
        if override is True:
            self.frames = {}
        out = {}
        for f in self._select(frames):
            if f not in self.frames:
                self.frames[f] = self._get_frame(self.trajectory_map[f], f, **kwargs)
            out[f] = self.frames[f]
        return out

    # ---- the batched analysis (replaces trajectory.py:350-586)
    def analysis(self, frames="all", ncpus=1, ncpus_analysis=1, override=False, modular=False, rebuild=False,  # noqa: ARG002, FBT002, PLR0913
                 swap_atoms=None, forcefield=None, **gpu_kwargs):
        """Analyse frames; results go to ``analysis_output`` exactly as in the reference.
        ``ncpus`` / ``ncpus_analysis`` are accepted and ignored (one GPU batch replaces the pools)."""
        todo = self._select(frames)
        if override is False and self.analysis_output:
            have = self.analysis_output
            have = dict.keys(have._load()) if isinstance(have, AnalysisOutput) else have.keys()
            todo = [f for f in todo if f not in have]
        if not todo:
            return
        rank, size = world()
        lo, hi = shard_range(len(todo), rank, size)
        mine = todo[lo:hi]
        if modular:
            records, keys = self._analyse_modular(mine, rebuild, swap_atoms, forcefield, gpu_kwargs)
        else:
            records, keys = self._analyse_single(mine, swap_atoms, forcefield, gpu_kwargs)
        if size > 1:
            records, keys = gather_records(records, keys)
        out = self.analysis_output
        if not isinstance(out, AnalysisOutput):          # a plain dict assigned by the user: keep its contents
            out = self.analysis_output = AnalysisOutput(out)
        out._add_block(records, keys)

    def _elements_of(self, frame, swap_atoms, forcefield):
        ms = self._get_frame(self.trajectory_map[frame], frame, swap_atoms=swap_atoms, forcefield=forcefield)
        if "elements" not in ms.system:
            raise U._FunctionError(
                "The 'elements' key is missing in the 'system' dictionary attribute of the MolecularSystem object. "
                "Which means, you need to decipher the forcefield based atom keys first (see manual).")
        return ms.system["elements"]

    def _decode_coordinates(self, frames):
        """[F, N, 3] float64 for the requested frames (atom order is frame invariant)."""
        out = None
        with self.filepath.open("rb") as fh, mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ) as mm:
            for k, f in enumerate(frames):
                start, end = self.trajectory_map[f]
                xyz = self._decode_xyz_only(mm[start:end])
                if out is None:
                    out = np.empty((len(frames),) + xyz.shape, dtype=np.float64)
                if xyz.shape != out.shape[1:]:
                    raise _TrajectoryError(f"frame {f} has {xyz.shape[0]} atoms, expected {out.shape[1]}")
                out[k] = xyz
        return out

    def _decode_xyz_only(self, raw):
        return np.asarray(self._decode_frame(raw)["coordinates"], dtype=np.float64)

    def _decode_into(self, frames, out, threads=0):
        """Decode ``frames`` into ``out[:len(frames)]`` ([cap, N, 3] float64, e.g. a pinned staging buffer).  The
        format classes override this with their native multi-threaded decoders writing in place."""
        out[:len(frames)] = self._decode_coordinates(frames)

    def _analyse_single(self, frames, swap_atoms, forcefield, gpu_kwargs):
        """Frames -> records through :class:`engine.FramePipeline`: the native decoder fills the pinned buffer of one
        slot on all host threads while the previous batch is on the GPU; one workspace / ray capacity for the whole
        trajectory; the only synchronisation per batch is the wait for its own records."""
        from .engine import host_ray_capacity
        eng = default_engine()
        if not frames:
            return np.zeros((0, eng.rec_stride)), RecordKeys.single([])
        elements = self._elements_of(frames[0], swap_atoms, forcefield)
        B = max(1, self.batch_frames)
        # a short first batch gets the GPU going after 1/8 of a batch's decode time; the rest are full batches.  A
        # frame list shorter than two batches (a rank's shard of a strongly scaled run) is cut in four, so that its
        # decode still overlaps the kernels
        n = len(frames)
        step = B if n >= 2 * B else min(B, max(2048, -(-n // 4)))
        head = min(step, max(256, step // 8)) if n > step else step
        chunks = [frames[:head]] + [frames[s:s + step] for s in range(head, n, step)]
        threads = _host_threads()
        pipe = eng.frame_pipeline(elements, B, slots=2, **gpu_kwargs)
        nslots = len(pipe.slots)
        records = np.empty((len(frames), eng.rec_stride), dtype=np.float64)
        starts = np.concatenate([[0], np.cumsum([len(c) for c in chunks])])
        # the records of a finished batch leave their pinned slot on a helper thread (the copy into fresh pageable
        # memory is page-fault bound, ~0.2 ms/MB, and would otherwise delay the decode of the next batch)
        from concurrent.futures import ThreadPoolExecutor
        copies = {}

        def take(j):
            src = pipe.collect(j % nslots)
            lo = int(starts[j])

            def job():
                records[lo:lo + len(chunks[j])] = src
            copies[j % nslots] = mover.submit(job)

        with ThreadPoolExecutor(max_workers=1) as mover:
            for k, chunk in enumerate(chunks):
                self._decode_into(chunk, pipe.buffer(k % nslots), threads)
                if k == 0:      # ray-list capacity for the whole trajectory from the first batch's bounding boxes
                    pipe.ensure_ray_capacity(host_ray_capacity(pipe.buffer(0)[:len(chunk)], pipe.vdw_max,
                                                               float(gpu_kwargs.get("adjust", 1.0))))
                if k % nslots in copies:                 # the slot's previous records must have left it
                    copies.pop(k % nslots).result()
                pipe.submit(k % nslots, len(chunk))
                if k + 1 >= nslots:                      # the slot the next batch will be decoded into must be free
                    take(k + 1 - nslots)
            for j in range(max(0, len(chunks) - nslots + 1), len(chunks)):
                take(j)
            for fut in copies.values():
                fut.result()
        records = self._redo_overflowed(records, frames, elements, gpu_kwargs)
        return records, RecordKeys.single(frames)

    def _redo_overflowed(self, records, frames, elements, gpu_kwargs):
        """Frames whose ray count exceeded the trajectory-wide capacity (a molecule that grew far beyond the first
        batch's bounding boxes) are re-analysed on their own, with a capacity sized for them."""
        over = np.nonzero(records[:, _lib.REC_STATUS].astype(np.int64) & _lib.ST_RAY_CAP)[0]
        if over.size:
            redo = [frames[i] for i in over]
            coords = self._decode_coordinates(redo)
            records[over] = default_engine().analyse_frames(coords, elements, **gpu_kwargs).host_records()
        return records

    def _analyse_modular(self, frames, rebuild, swap_atoms, forcefield, gpu_kwargs):
        """modular=True: every frame is decomposed into discrete molecules on the GPU (rebuilt across
        the periodic boundary if asked) and every molecule analysed, whole batches of frames at once."""
        from .periodic import analyse_cells
        eng = default_engine()
        if not frames:
            return np.zeros((0, eng.rec_stride)), RecordKeys.single([])
        elements = self._elements_of(frames[0], swap_atoms, forcefield)
        recs, kf, km = [], [], []
        # cells per batch: the flood fill runs one warp per cell and is bound by memory latency, so it wants
        # thousands of cells in flight; batch k+1 is decoded on a helper thread (the C++ decoder and the GPU calls
        # both release the GIL) while batch k is on the GPU
        step = max(1, min(self.modular_batch_cells, (1 << 24) // max(1, len(elements))))
        chunks = [frames[s:s + step] for s in range(0, len(frames), step)]
        from concurrent.futures import ThreadPoolExecutor
        N = len(elements)
        slots = [eng.pinned(("cells", i), (step, N, 3)) for i in range(2)]      # decoded straight into pinned memory

        def decode(k):
            dst = slots[k % 2][:len(chunks[k])]
            xyz, lat = self._decode_cells(chunks[k], out=dst.numpy())
            # (a format class without an in-place native decoder returns its own array)
            return (dst if np.shares_memory(xyz, dst.numpy()) else xyz), lat

        with ThreadPoolExecutor(max_workers=1) as pool:
            pending = pool.submit(decode, 0)
            for k, chunk in enumerate(chunks):
                coords, lattice = pending.result()
                if k + 1 < len(chunks):
                    pending = pool.submit(decode, k + 1)
                if rebuild and lattice is None:
                    raise _TrajectoryError("rebuild=True needs periodic frames (no lattice / CRYST1 record found)")
                r, (fi, mi), _ = analyse_cells(coords, elements, lattice=lattice, rebuild=rebuild, engine=eng,
                                               key_arrays=True, **gpu_kwargs)
                recs.append(r)
                kf.append(np.asarray(chunk, dtype=np.int64)[fi])
                km.append(mi)
        return np.concatenate(recs), RecordKeys(np.concatenate(kf), np.concatenate(km))

    def _decode_cells(self, frames, out=None):  # noqa: ARG002  (format classes with a native decoder fill `out`)
        """([F,N,3] coordinates, [F,3,3] lattices | None) for the requested frames."""
        coords, lats = None, []
        with self.filepath.open("rb") as fh, mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ) as mm:
            for k, f in enumerate(frames):
                start, end = self.trajectory_map[f]
                fr = self._decode_frame(mm[start:end])
                xyz = np.asarray(fr["coordinates"], dtype=np.float64)
                if coords is None:
                    coords = np.empty((len(frames),) + xyz.shape, dtype=np.float64)
                if xyz.shape != coords.shape[1:]:
                    raise _TrajectoryError(f"frame {f} has {xyz.shape[0]} atoms, expected {coords.shape[1]}")
                coords[k] = xyz
                if "lattice" in fr:
                    lats.append(np.asarray(fr["lattice"], dtype=np.float64))
        lattice = np.stack(lats) if len(lats) == len(frames) else None
        return coords, lattice

    # ---- output (trajectory.py:251-347)
    def save_analysis(self, filepath=None, **kwargs):
        if filepath is None:
            filepath = pathlib.Path.cwd() / f"{self.system_id}_pywindow_analysis"
        filepath = pathlib.Path(filepath)
        if filepath.suffix != ".json":
            filepath = pathlib.Path(str(filepath) + ".json")
        out = {str(f): {str(k): v for k, v in dict(mols.items()).items()} for f, mols in self.analysis_output.items()}
        with filepath.open("w") as fh:
            json.dump(out, fh, default=U.to_list, **kwargs)

    def save_frames(self, frames, filepath=None, filetype="pdb", **kwargs):
        if filepath is None:
            filepath = pathlib.Path.cwd() / str(self.system_id)
        for f, ms in self.get_frames(frames).items():
            ms.dump_system(filepath=pathlib.Path(f"{filepath}_{f}.{filetype}"), **kwargs)


def _float_block(lines):
    return np.array(b" ".join(lines).split(), dtype=np.float64)


class DLPOLY(Trajectory):
    """DL_POLY_C HISTORY file.  The file is indexed once (byte offsets of every ``timestep`` record);
    frames are decoded on demand from the memory map."""

    _imcon = {0: "nonperiodic", 1: "cubic", 2: "orthorhombic", 3: "parallelepiped", 4: "truncated octahedral",
              5: "rhombic dodecahedral", 6: "x-y parallelogram", 7: "hexagonal prism"}
    _keytrj = {0: "coordinates", 1: "coordinates and velocities", 2: "coordinates, velocities and forces"}

    def __init__(self, filepath):
        super().__init__(filepath)
        self._map_history()

    def _map_history(self):
        lib = _lib.load()
        if self.filepath.stat().st_size == 0:
            raise _TrajectoryError("no 'timestep' record found: not a DL_POLY HISTORY file")
        mm, buf = self._mapped()
        size = len(mm)
        cap = max(1024, size // 2048)
        while True:
            starts = np.empty(cap, dtype=np.int64)
            n = int(lib.pwb_history_index(buf.ctypes.data, size, starts.ctypes.data, cap))
            if n <= cap:
                break
            cap = n
        starts = starts[:n]
        if not n:
            raise _TrajectoryError("no 'timestep' record found: not a DL_POLY HISTORY file")
        s0 = int(starts[0])
        head = bytes(mm[:s0]).decode("utf-8").split("\n")
        first = bytes(mm[s0:mm.find(b"\n", s0)]).split()
        header = [int(i) for i in head[1].split()]
        pairs = np.empty((n, 2), dtype=np.int64)
        pairs[:, 0] = starts
        pairs[:-1, 1] = starts[1:]
        pairs[-1, 1] = size
        self.trajectory_map = FrameMap(pairs, header=header)
        self.content_type = self._keytrj[header[0]]
        self.periodic_boundary = self._imcon[header[1]]
        self.no_of_atoms = header[2]
        self._frame_keytrj, self._frame_imcon = int(first[3]), int(first[4])
        self.no_of_frames = n
        self.check_log = ""

    def _native_decode(self, frames, want_lattice, out=None, threads=0):
        """[F,N,3] (and [F,3,3] lattices) through the multi-threaded C++ decoder (pwb_history_decode); ``out``:
        write the coordinates into ``out[:F]`` (a pinned staging buffer) instead of a new array."""
        lib = _lib.load()
        n, N = len(frames), self.no_of_atoms
        starts, ends = _bounds(self, frames)
        if out is None:
            xyz = np.empty((n, N, 3), dtype=np.float64)
        else:
            if out.shape[1:] != (N, 3) or out.shape[0] < n or out.dtype != np.float64 or not out.flags.c_contiguous:
                raise _TrajectoryError(f"decode buffer {out.shape} does not fit {n} frames of {N} atoms")
            xyz = out
        periodic = self._frame_imcon in (1, 2, 3)
        lat = np.empty((n, 3, 3), dtype=np.float64) if (want_lattice and periodic) else None
        mm, buf = self._mapped()
        bad = int(lib.pwb_history_decode(buf.ctypes.data, len(mm), starts.ctypes.data, ends.ctypes.data, n, N,
                                         self._frame_keytrj, self._frame_imcon, xyz.ctypes.data,
                                         lat.ctypes.data if lat is not None else None, threads))
        if bad:
            raise _TrajectoryError(f"frame {frames[bad - 1]} is malformed (expected {N} atoms, keytrj "
                                   f"{self._frame_keytrj}, imcon {self._frame_imcon})")
        return xyz, lat

    def _decode_coordinates(self, frames):
        return self._native_decode(frames, False)[0]

    def _decode_into(self, frames, out, threads=0):
        self._native_decode(frames, False, out=out, threads=threads)

    def _decode_cells(self, frames, out=None):
        return self._native_decode(frames, True, out=out)

    def _layout(self, lines):
        info = lines[0].split()
        keytrj, imcon = int(info[3]), int(info[4])
        first = 4 if imcon in (1, 2, 3) else 1
        return info, keytrj, imcon, first, keytrj + 2

    def _decode_frame(self, raw):
        lines = [ln for ln in bytes(raw).split(b"\n") if ln.strip()]
        info, keytrj, imcon, first, stride = self._layout(lines)
        frame = {"frame_info": {"nstep": int(info[1]), "natms": int(info[2]), "keytrj": keytrj, "imcon": imcon,
                                "tstep": float(info[5])}}
        if first == 4:
            frame["lattice"] = _float_block(lines[1:4]).reshape(3, 3).T
            frame["unit_cell"] = U.lattice_array_to_unit_cell(frame["lattice"])
        body = lines[first:]
        frame["atom_ids"] = np.array([ln.split()[0].decode("utf-8") for ln in body[0::stride]])
        frame["coordinates"] = _float_block(body[1::stride]).reshape(-1, 3)
        if keytrj >= 1:
            frame["velocities"] = _float_block(body[2::stride]).reshape(-1, 3)
        if keytrj >= 2:
            frame["forces"] = _float_block(body[3::stride]).reshape(-1, 3)
        return frame

    def _decode_xyz_only(self, raw):
        lines = [ln for ln in bytes(raw).split(b"\n") if ln.strip()]
        _, _, _, first, stride = self._layout(lines)
        return _float_block(lines[first + 1::stride]).reshape(-1, 3)


def _native_xyz_like(traj, entry, frames, natoms, out=None, threads=0):
    """Shared driver of pwb_xyz_decode / pwb_pdb_decode: frames -> [F, natoms, 3] float64 (into ``out[:F]`` if given)."""
    lib = _lib.load()
    n = len(frames)
    starts, ends = _bounds(traj, frames)
    if out is None:
        xyz = np.empty((n, natoms, 3), dtype=np.float64)
    else:
        if out.shape[1:] != (natoms, 3) or out.shape[0] < n or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise _TrajectoryError(f"decode buffer {out.shape} does not fit {n} frames of {natoms} atoms")
        xyz = out
    mm, buf = traj._mapped()
    bad = int(getattr(lib, entry)(buf.ctypes.data, len(mm), starts.ctypes.data, ends.ctypes.data, n, natoms,
                                  xyz.ctypes.data, threads))
    if bad:
        raise _TrajectoryError(f"frame {frames[bad - 1]} is malformed (expected {natoms} atoms)")
    return xyz



class XYZ(Trajectory):
    """Multi-frame XYZ file: every frame is ``natoms`` / comment / natoms atom lines."""

    def __init__(self, filepath):
        super().__init__(filepath)
        # frame map by the native indexer (pwb_xyz_index): byte range and atom count of every frame
        lib = _lib.load()
        self._natoms = np.zeros(0, dtype=np.int32)
        if self.filepath.stat().st_size:
            with self.filepath.open("rb") as fh, mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ) as mm:
                buf = np.frombuffer(mm, dtype=np.uint8)
                cap = 1 << 16
                while True:
                    starts, ends = np.empty(cap, dtype=np.int64), np.empty(cap, dtype=np.int64)
                    natoms = np.empty(cap, dtype=np.int32)
                    n = int(lib.pwb_xyz_index(buf.ctypes.data, len(mm), starts.ctypes.data, ends.ctypes.data,
                                              natoms.ctypes.data, cap))
                    if n <= cap:
                        break
                    cap = n
                del buf
            if n < 0:
                raise _FormatError(f"expected an atom count at byte {-n - 1} of {self.filepath}")
            for k in range(n):
                self.trajectory_map[k] = [int(starts[k]), int(ends[k])]
            self._natoms = natoms[:n].copy()
        self.no_of_frames = len(self.trajectory_map)

    def _decode_coordinates(self, frames):
        """[F,N,3] through the multi-threaded C++ decoder (pwb_xyz_decode) when the frames share one atom count."""
        counts = {int(self._natoms[f]) for f in frames}
        if len(counts) != 1:
            return super()._decode_coordinates(frames)
        return _native_xyz_like(self, "pwb_xyz_decode", frames, counts.pop())

    def _decode_into(self, frames, out, threads=0):
        counts = {int(self._natoms[f]) for f in frames}
        if len(counts) != 1:
            return super()._decode_into(frames, out, threads)
        _native_xyz_like(self, "pwb_xyz_decode", frames, counts.pop(), out=out, threads=threads)

    def _decode_frame(self, raw):
        lines = bytes(raw).decode("utf-8").split("\n")
        natoms = int(lines[0].split()[0])
        tokens = [ln.split() for ln in lines[2:2 + natoms]]
        return {"remarks": lines[1], "atom_ids": np.array([t[0] for t in tokens]),
                "elements": np.array([t[0] for t in tokens]),
                "coordinates": np.array([[float(t[1]), float(t[2]), float(t[3])] for t in tokens])}


class PDB(Trajectory):
    """Multi-model PDB file: frames end with an ``END`` record."""

    def __init__(self, filepath):
        super().__init__(filepath)
        with self.filepath.open("rb") as fh, mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ) as mm:
            start, k, size = 0, 0, len(mm)
            # whole-line END / ENDMDL records only (the reference splits on lines that are exactly "END",
            # trajectory.py:960-975); a slice is a frame only if it holds an atom record, so the closing END
            # after the last ENDMDL of a multi-model file does not become an empty frame
            atom_rec = re.compile(rb"^(?:ATOM  |HETATM)", flags=re.MULTILINE)
            for m in re.finditer(rb"^(?:ENDMDL|END)[ \t\r]*$\n?", mm, flags=re.MULTILINE):
                if atom_rec.search(mm, start, m.start()):
                    self.trajectory_map[k] = [start, m.end()]
                    k += 1
                start = m.end()
            if start < size and atom_rec.search(mm, start, size):
                self.trajectory_map[k] = [start, size]
                k += 1
        self.no_of_frames = k

    def _decode_coordinates(self, frames):
        """[F,N,3] through the multi-threaded C++ decoder (pwb_pdb_decode); the atom count is the first frame's."""
        if not frames:
            return super()._decode_coordinates(frames)
        with self.filepath.open("rb") as fh, mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ) as mm:
            start, end = self.trajectory_map[frames[0]]
            natoms = len(self._decode_frame(mm[start:end])["coordinates"])
        return _native_xyz_like(self, "pwb_pdb_decode", frames, natoms)

    def _decode_into(self, frames, out, threads=0):
        _native_xyz_like(self, "pwb_pdb_decode", frames, out.shape[1], out=out, threads=threads)

    def _decode_frame(self, raw):
        lines = bytes(raw).decode("utf-8").split("\n")
        atoms = [ln for ln in lines if ln[:6] in ("HETATM", "ATOM  ")]
        frame = {"remarks": [ln for ln in lines if ln[:6] == "REMARK"]}
        cell = [float(ln[a:b]) for ln in lines if ln[:6] == "CRYST1"
                for a, b in ((6, 15), (15, 24), (24, 33), (33, 40), (40, 47), (47, 54))]
        if cell:
            frame["unit_cell"] = np.array(cell)
            frame["lattice"] = U.unit_cell_to_lattice_array(frame["unit_cell"])
        frame["atom_ids"] = np.array([ln[12:16].strip() for ln in atoms], dtype="<U8")
        frame["elements"] = np.array([ln[76:78].strip() for ln in atoms], dtype="<U8")
        frame["coordinates"] = np.array([[float(ln[30:38]), float(ln[38:46]), float(ln[46:54])] for ln in atoms])
        return frame
