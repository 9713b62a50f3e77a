"""Batched GPU driver for the pore-geometry path.

``Engine`` owns one ``pwb_ctx`` (device context) and turns batches of molecule-frames into result
records through the C ABI (``include/pwb.h``).  torch is used only for device memory and streams.
There is no CPU fallback: without a CUDA device or without ``libpwb.so`` construction raises.
"""
from __future__ import annotations

import ctypes
import logging
import warnings

import numpy as np
import torch

from . import _lib
from .tables import atomic_mass, atomic_vdw_radius

logger = logging.getLogger("pywindow")

DEFAULT_WIN_CAP = 16


class PwbNoDeviceError(RuntimeError):
    pass


def element_types(elements):
    """Map element symbols to small integer types + per-type vdW radius / mass tables
    (replaces the per-atom dict lookups of utilities.py:96-148, :358-364, :382)."""
    up = [str(e).upper() for e in elements]
    uniq = sorted(set(up))
    if len(uniq) > _lib.MAX_TYPES:
        raise ValueError(f"more than {_lib.MAX_TYPES} distinct elements in one batch")
    idx = {e: i for i, e in enumerate(uniq)}
    try:
        vdw = np.array([atomic_vdw_radius[e] for e in uniq], dtype=np.float64)
        mass = np.array([atomic_mass[e] for e in uniq], dtype=np.float64)
    except KeyError as exc:
        raise KeyError(f"unknown element {exc.args[0]!r}") from None
    return np.array([idx[e] for e in up], dtype=np.int32), vdw, mass


class BatchResult:
    """Records of one batch plus views of the per-stage lists in the workspace."""

    def __init__(self, records, workspace, layout, n_items, total_atoms, ray_cap, win_cap):
        self.records = records            # torch [F, rec_stride] float64 (device)
        self._ws = workspace
        self._layout = layout
        self.n_items, self.total_atoms, self.ray_cap, self.win_cap = n_items, total_atoms, ray_cap, win_cap

    def _view(self, off, dtype, shape):
        n = int(np.prod(shape))
        item = torch.empty((), dtype=dtype).element_size()
        return self._ws[off: off + n * item].view(dtype).view(*shape)

    @property
    def clear_k(self):
        return self._view(self._layout.clear_k, torch.int32, (self.n_items, self.ray_cap))

    @property
    def clear_pos(self):
        return self._view(self._layout.clear_pos, torch.int32, (self.n_items, self.ray_cap))

    @property
    def clear_a2(self):
        return self._view(self._layout.clear_a2, torch.float64, (self.n_items, self.ray_cap))

    @property
    def labels(self):
        return self._view(self._layout.labels, torch.int32, (self.n_items, self.ray_cap))

    @property
    def win_k(self):
        return self._view(self._layout.win_k, torch.int32, (self.n_items, self.win_cap))

    @property
    def item_d(self):
        return self._view(self._layout.item_d, torch.float64, (self.n_items, 16))

    @property
    def item_i(self):
        return self._view(self._layout.item_i, torch.int32, (self.n_items, 32))

    def host_records(self):
        return self.records.cpu().numpy()


class Engine:
    def __init__(self, device=None, table_cap=4100, win_cap=DEFAULT_WIN_CAP, chunks=2):
        if not torch.cuda.is_available():
            raise PwbNoDeviceError("pywindow_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.win_cap = win_cap
        self.table_cap = table_cap
        self._ctx = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pwb_ctx_create(ctypes.byref(self._ctx), self.device.index, table_cap))
        self.set_chunks(chunks)
        self._pinned = {}

    def set_chunks(self, chunks):
        """Sub-batches pipelined over internal CUDA streams per call (1 = strictly sequential stages)."""
        self.chunks = int(chunks)
        self.lib.pwb_ctx_set_chunks(self._ctx, self.chunks)

    def set_option(self, option, value):
        """Test switches of the library (``_lib.OPTION_*``, include/pwb.h)."""
        _lib.check(self.lib.pwb_ctx_set_option(self._ctx, int(option), int(value)))

    def close(self):
        if self._ctx:
            self.lib.pwb_ctx_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(self.lib.pwb_launch_count(self._ctx))

    @property
    def rec_stride(self):
        return _lib.REC_WIN0 + 4 * self.win_cap

    # ------------------------------------------------------------------ batch construction
    def _dev(self, a, dtype):
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(self.device)

    def _make_batch(self, coords, item_off, atom_type, type_vdw, type_mass, shared_types, max_atoms,
                    centre_override, stages, adjust, increment, increment2, pore_opt, ray_cap,
                    opt_start=None, opt_bounds=None):
        F = int(item_off.numel() - 1)
        total = int(coords.shape[0])
        b = _lib.Batch()
        b.n_items, b.max_atoms, b.n_types = F, int(max_atoms), int(type_vdw.numel())
        b.shared_types, b.win_cap, b.rec_stride = int(bool(shared_types)), self.win_cap, self.rec_stride
        b.stages = int(stages)
        b.adjust, b.increment, b.increment2 = float(adjust), float(increment), float(increment2)
        b.pore_opt, b.total_atoms = int(bool(pore_opt)), total
        b.coords, b.item_off, b.atom_type = coords.data_ptr(), item_off.data_ptr(), atom_type.data_ptr()
        b.type_vdw, b.type_mass = type_vdw.data_ptr(), type_mass.data_ptr()
        b.centre_override = centre_override.data_ptr() if centre_override is not None else None
        b.opt_start = opt_start.data_ptr() if opt_start is not None else None
        b.opt_bounds = opt_bounds.data_ptr() if opt_bounds is not None else None
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        if ray_cap is None:
            cap = ctypes.c_int(0)
            b.ray_cap = 0
            _lib.check(self.lib.pwb_ray_capacity(self._ctx, ctypes.byref(b), stream, ctypes.byref(cap)))
            ray_cap = max(64, (cap.value + 63) // 64 * 64)
        b.ray_cap = int(ray_cap)
        return b, stream, F, total

    def analyse(self, coords, item_off, atom_type, type_vdw, type_mass, *, shared_types, max_atoms,
                centre_override=None, stages=_lib.STAGE_ALL, adjust=1.0, increment=1.0, increment2=0.1,
                pore_opt=True, ray_cap=None, opt_start=None, opt_bounds=None):
        """Run the stages on a batch whose inputs are (or are moved to) device tensors."""
        if int(item_off.shape[0] if hasattr(item_off, "shape") else len(item_off)) <= 1:      # empty batch
            empty = torch.zeros((0, self.rec_stride), dtype=torch.float64, device=self.device)
            return BatchResult(empty, torch.zeros(0, dtype=torch.uint8, device=self.device), _lib.WsLayout(), 0, 0,
                               int(ray_cap or 64), self.win_cap)
        with torch.cuda.device(self.device):
            coords = self._dev(coords, torch.float64).view(-1, 3)
            item_off = self._dev(item_off, torch.int32)
            atom_type = self._dev(atom_type, torch.int32)
            type_vdw = self._dev(type_vdw, torch.float64)
            type_mass = self._dev(type_mass, torch.float64)
            if centre_override is not None:
                centre_override = self._dev(centre_override, torch.float64).view(-1, 3)
            if opt_start is not None:
                opt_start = self._dev(opt_start, torch.float64).view(-1, 3)
            if opt_bounds is not None:
                opt_bounds = self._dev(opt_bounds, torch.float64).view(-1, 6)
            n_items = int(item_off.numel() - 1)
            for name, t in (("centre_override", centre_override), ("opt_start", opt_start), ("opt_bounds", opt_bounds)):
                if t is not None and t.shape[0] != n_items:
                    raise ValueError(f"{name} needs one row per item ({n_items}), got {t.shape[0]}")
            b, stream, F, total = self._make_batch(coords, item_off, atom_type, type_vdw, type_mass,
                                                   shared_types, max_atoms, centre_override, stages, adjust,
                                                   increment, increment2, pore_opt, ray_cap, opt_start, opt_bounds)
            layout = _lib.WsLayout()
            self.lib.pwb_workspace_layout(F, total, int(max_atoms), b.ray_cap, self.win_cap, ctypes.byref(layout))
            ws = torch.empty(layout.total, dtype=torch.uint8, device=self.device)
            rec = torch.empty((F, self.rec_stride), dtype=torch.float64, device=self.device)
            b.records, b.workspace, b.workspace_bytes = rec.data_ptr(), ws.data_ptr(), layout.total
            _lib.check(self.lib.pwb_full_analysis_batch(self._ctx, ctypes.byref(b), stream))
            res = BatchResult(rec, ws, layout, F, total, b.ray_cap, self.win_cap)
            res._keep = (coords, item_off, atom_type, type_vdw, type_mass, centre_override, opt_start, opt_bounds)
            res._batch, res._stream = b, stream          # the descriptor, for stage-wise re-runs on the same workspace
            return res

    def analyse_ray_sharded(self, elements, coords, *, adjust=1.0, increment=1.0, increment2=0.1, pore_opt=True,
                            centre_override=None, stages=_lib.STAGE_ALL):
        """ONE molecule with a dense ray set (BASELINE config 5: adjust ~ 100 -> ~1e5 rays), its rays sharded over
        the ranks of the initialised ``torch.distributed`` group: every rank runs the cheap scalar /
        pore-centre stages on the replicated molecule, analyses only its own contiguous range of ray tiles
        (pre-test + path clearance are independent per ray, utilities.py:1441-1470), the clear-ray lists are
        all-gathered in rank order (= ascending ray index: 16 B per clear ray over NCCL / NVLink), and DBSCAN + window
        refinement run on the merged list on every rank, so all ranks return the same record.  Collective: every
        rank must call it with the same arguments.  With one rank it is ``analyse_molecules`` on one item."""
        import torch.distributed as dist
        from .distributed import world
        rank, size = world()
        el = list(elements)
        X = np.ascontiguousarray(coords, dtype=np.float64).reshape(-1, 3)
        kw = dict(adjust=adjust, increment=increment, increment2=increment2, pore_opt=pore_opt)
        if centre_override is not None:
            kw["centre_override"] = np.asarray(centre_override, dtype=np.float64).reshape(1, 3)
        if size == 1:
            return self.analyse_molecules([(el, X)], stages=stages, **kw)
        pre = stages & (_lib.STAGE_SCALARS | _lib.STAGE_AVGD | _lib.STAGE_OPT)
        res = self.analyse_molecules([(el, X)], stages=pre | _lib.STAGE_SCALARS, **kw)
        if not stages & _lib.STAGE_RAYS:
            return res
        b, stream = res._batch, res._stream
        with torch.cuda.device(self.device):
            # window rays: P = int(log10(4 pi R^2) 250 adjust), R = max_dim / 2 (utilities.py:1398-1409); the tile
            # ranges only have to PARTITION the tiles (the kernel clamps them to its own P), so a host-side P that is
            # off by one at an integer boundary is harmless: the last rank's range is open ended
            D = float(res.records[0, _lib.REC_MAXD].item())
            P = max(2, int(np.log10(4.0 * np.pi * (D / 2) ** 2) * 250.0 * float(adjust)))
            tiles = -(-P // _lib.RAY_TILE)
            per = -(-tiles // size)
            b.stages = _lib.STAGE_RAYS
            b.ray_tile_begin = rank * per
            b.ray_tile_end = (rank + 1) * per if rank + 1 < size else (1 << 20)
            _lib.check(self.lib.pwb_full_analysis_batch(self._ctx, ctypes.byref(b), stream))
            ii = res.item_i
            mine = ii[0, 2:4].clone()                                   # I_NSURV, I_NCLEAR of this rank's range
            counts = [torch.zeros_like(mine) for _ in range(size)]
            dist.all_gather(counts, mine)
            counts = torch.stack(counts).cpu().numpy()
            n_each, n_clear, n_surv = counts[:, 1], int(counts[:, 1].sum()), int(counts[:, 0].sum())
            kmax = int(n_each.max())
            if kmax:
                # one payload row per clear ray: ray index, arg-min step, 2*min clearance
                pay = torch.zeros((kmax, 3), dtype=torch.float64, device=self.device)
                n = int(n_each[rank])
                pay[:n, 0] = res.clear_k[0, :n].to(torch.float64)
                pay[:n, 1] = res.clear_pos[0, :n].to(torch.float64)
                pay[:n, 2] = res.clear_a2[0, :n]
                out = torch.empty((size * kmax, 3), dtype=torch.float64, device=self.device)
                dist.all_gather_into_tensor(out, pay)
                out = out.view(size, kmax, 3)
                merged = torch.cat([out[r, :int(n_each[r])] for r in range(size)])
                res.clear_k[0, :n_clear] = merged[:, 0].to(torch.int32)
                res.clear_pos[0, :n_clear] = merged[:, 1].to(torch.int32)
                res.clear_a2[0, :n_clear] = merged[:, 2]
            ii[0, 2], ii[0, 3] = n_surv, n_clear
            res.records[0, _lib.REC_NCLEAR], res.records[0, _lib.REC_NSURV] = float(n_clear), float(n_surv)
            if n_clear == 0:
                ii[0, 0] |= _lib.ST_NO_CLEAR_RAY
                res.records[0, _lib.REC_STATUS] = float(int(ii[0, 0]))
            b.ray_tile_begin = b.ray_tile_end = 0
            b.stages = stages & (_lib.STAGE_CLUSTER | _lib.STAGE_WINDOWS)
            if b.stages:
                _lib.check(self.lib.pwb_full_analysis_batch(self._ctx, ctypes.byref(b), stream))
        return res

    def analyse_frames(self, coords, elements, **kw):
        """``coords``: [F, N, 3] (numpy or torch), one molecule with atom order ``elements``."""
        F, N = int(coords.shape[0]), int(coords.shape[1])
        types, vdw, mass = element_types(elements)
        off = torch.arange(0, (F + 1) * N, N, dtype=torch.int32)
        flat = coords.reshape(F * N, 3)
        return self.analyse(flat, off, types, vdw, mass, shared_types=True, max_atoms=N, **kw)

    def analyse_molecules(self, mols, **kw):
        """``mols``: list of (elements, coords[N_i,3]) with arbitrary sizes / compositions."""
        all_el = [e for el, _ in mols for e in el]
        types, vdw, mass = element_types(all_el)
        sizes = [len(el) for el, _ in mols]
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        flat = np.concatenate([np.asarray(c, dtype=np.float64).reshape(-1, 3) for _, c in mols])
        return self.analyse(flat, off, types, vdw, mass, shared_types=False, max_atoms=max(sizes), **kw)

    def pinned(self, key, shape, dtype=torch.float64):
        """Pooled pinned host tensor for ``key`` (grown, never shrunk): page-locking a few hundred MB costs more than
        analysing the frames they hold, and pageable copies run at a fifth of the PCIe rate."""
        k = ("pinned",) + tuple(key)
        t = self._pinned.get(k)
        need = int(np.prod(shape))
        if t is None or t.numel() < need or t.dtype != dtype:
            t = self._pinned[k] = torch.empty(max(need, 1), dtype=dtype, pin_memory=True)
        return t[:need].view(*shape)

    def frame_pipeline(self, elements, cap_frames, ray_cap=0, slots=2, **params):
        """:class:`FramePipeline` bound to this engine (the trajectory front-ends' entry point).  Pipelines are
        pooled: a later trajectory of the same molecule, batch size and parameters reuses the pinned / device buffers
        (pinning a few hundred MB costs more than analysing 100 000 frames)."""
        key = (tuple(str(e) for e in elements), int(cap_frames), int(slots), tuple(sorted(params.items())))
        pipe = self._pinned.get(key)
        if pipe is None or any(sl["busy"] for sl in pipe.slots):
            if len(self._pinned) >= 4:
                self._pinned.pop(next(iter(self._pinned)))
            pipe = self._pinned[key] = FramePipeline(self, elements, cap_frames, 0, slots=slots, **params)
        if ray_cap:
            pipe.ensure_ray_capacity(ray_cap)
        return pipe

    # ------------------------------------------------------------------ host-buffer path (e2e)
    def analyse_frames_host(self, coords_host, elements, *, ray_cap, out=None, stages=_lib.STAGE_ALL,
                            adjust=1.0, increment=1.0, increment2=0.1, pore_opt=True, state=None, wait=True):
        """Host numpy/pinned [F,N,3] in, host records out: H2D + kernels + D2H inside the call
        (``pwb_full_analysis_host``).  ``state`` caches device staging buffers between calls.
        ``wait=False`` (``pwb_full_analysis_host_async``) returns at once; the pinned records are valid after
        the current torch stream has been synchronised -- alternate two ``state`` dicts on two streams to
        overlap one batch's copies with the other's kernels."""
        F, N = int(coords_host.shape[0]), int(coords_host.shape[1])
        with torch.cuda.device(self.device):
            st = state if state is not None else {}
            key = (F, N, ray_cap, self.win_cap, self.rec_stride, tuple(str(e) for e in elements))
            if st.get("key") != key:
                types, vdw, mass = element_types(elements)
                st.clear()
                st["key"] = key
                st["types"] = self._dev(types, torch.int32)
                st["vdw"] = self._dev(vdw, torch.float64)
                st["mass"] = self._dev(mass, torch.float64)
                st["off"] = torch.arange(0, (F + 1) * N, N, dtype=torch.int32, device=self.device)
                st["d_coords"] = torch.empty((F * N, 3), dtype=torch.float64, device=self.device)
                st["d_rec"] = torch.empty((F, self.rec_stride), dtype=torch.float64, device=self.device)
                layout = _lib.WsLayout()
                self.lib.pwb_workspace_layout(F, F * N, N, ray_cap, self.win_cap, ctypes.byref(layout))
                st["layout"] = layout
                st["ws"] = torch.empty(layout.total, dtype=torch.uint8, device=self.device)
                st["h_rec"] = torch.empty((F, self.rec_stride), dtype=torch.float64).pin_memory()
            if isinstance(coords_host, torch.Tensor):
                h_ptr = coords_host.data_ptr()
            else:
                coords_host = np.ascontiguousarray(coords_host, dtype=np.float64)
                h_ptr = coords_host.ctypes.data
            h_rec = st["h_rec"] if out is None else out
            b, stream, _, _ = self._make_batch(st["d_coords"], st["off"], st["types"], st["vdw"], st["mass"],
                                               True, N, None, stages, adjust, increment, increment2, pore_opt,
                                               ray_cap)
            b.workspace, b.workspace_bytes = st["ws"].data_ptr(), st["layout"].total
            entry = self.lib.pwb_full_analysis_host if wait else self.lib.pwb_full_analysis_host_async
            if not wait:
                st["h_in"] = coords_host          # keep the source alive until the copy has run
            _lib.check(entry(self._ctx, ctypes.byref(b), h_ptr, h_rec.data_ptr(),
                              st["d_coords"].data_ptr(), st["d_rec"].data_ptr(), F * N, stream))
            return h_rec


def host_ray_capacity(coords, vdw_max, adjust=1.0, margin=64, sample=512):
    """Upper bound of the ray count the frames ``coords`` [F,N,3] can need, from the bounding boxes of an evenly
    spaced sample of them (the host-side twin of ``pwb_ray_capacity``; no device synchronisation).
    ``int(log10(4 pi D^2) 250 adjust)`` grows by 22 rays per 10 % of D, so ``margin`` rays cover frames ~30 % larger
    than any sampled; a frame that still overflows is flagged PWB_ST_RAY_CAP and re-analysed by the caller."""
    c = np.asarray(coords)
    if c.size == 0:
        return 64 + margin
    if c.shape[0] > sample:
        c = c[:: -(-c.shape[0] // sample)]
    c = np.ascontiguousarray(c.transpose(0, 2, 1))          # [F, 3, N]: the reductions run over contiguous rows
    ext = c.max(axis=2) - c.min(axis=2)
    D = float(np.sqrt((ext * ext).sum(axis=1)).max()) + 2.0 * float(vdw_max)
    P = int(np.log10(4.0 * np.pi * D * D) * 250.0 * float(adjust)) + 1
    return max(64, (P + margin + 63) // 64 * 64)


class FramePipeline:
    """Double-buffered host -> GPU -> host path for a trajectory of one molecule (the batched replacement of the
    per-frame loop of trajectory.py:473-494): ``slots`` pinned input buffers [cap_frames, N, 3] that the trajectory
    decoders fill directly, persistent device staging / workspace / record buffers per slot, and
    ``pwb_full_analysis_host_async`` so that batch i runs on the GPU while the host decodes batch i+1.  Nothing is
    allocated and nothing synchronises per batch except the wait for a slot's own records.  The ray-list workspace
    is sized by :meth:`ensure_ray_capacity` once the first frames are known."""

    def __init__(self, engine, elements, cap_frames, ray_cap=0, slots=2, stages=_lib.STAGE_ALL, adjust=1.0,
                 increment=1.0, increment2=0.1, pore_opt=True):
        self.eng, self.cap, self.N = engine, int(cap_frames), len(elements)
        self.ray_cap = 0
        self.params = dict(stages=stages, adjust=adjust, increment=increment, increment2=increment2, pore_opt=pore_opt)
        eng, dev, F, N = engine, engine.device, self.cap, self.N
        types, vdw, mass = element_types(elements)
        self.vdw_max = float(vdw.max())
        with torch.cuda.device(dev):
            self.types = eng._dev(types, torch.int32)
            self.vdw, self.mass = eng._dev(vdw, torch.float64), eng._dev(mass, torch.float64)
            self.off = torch.arange(0, (F + 1) * N, N, dtype=torch.int32, device=dev)
            self.layout = _lib.WsLayout()
            self.slots = []
            for _ in range(slots):
                self.slots.append({
                    "h_in": torch.empty((F, N, 3), dtype=torch.float64, pin_memory=True),
                    "h_rec": torch.empty((F, eng.rec_stride), dtype=torch.float64, pin_memory=True),
                    "d_in": torch.empty((F * N, 3), dtype=torch.float64, device=dev),
                    "d_rec": torch.empty((F, eng.rec_stride), dtype=torch.float64, device=dev),
                    "ws": None, "done": torch.cuda.Event(), "n": 0, "busy": False})
            for sl in self.slots:
                sl["in_np"], sl["rec_np"] = sl["h_in"].numpy(), sl["h_rec"].numpy()
        if ray_cap:
            self.ensure_ray_capacity(ray_cap)

    def ensure_ray_capacity(self, ray_cap):
        """(Re)allocate the per-slot workspaces for ray lists of ``ray_cap`` entries; no-op if already large enough.
        Must not be called while a batch is in flight."""
        if int(ray_cap) <= self.ray_cap:
            return
        if any(sl["busy"] for sl in self.slots):
            raise RuntimeError("ensure_ray_capacity while a batch is in flight")
        eng = self.eng
        self.ray_cap = int(ray_cap)
        with torch.cuda.device(eng.device):
            eng.lib.pwb_workspace_layout(self.cap, self.cap * self.N, self.N, self.ray_cap, eng.win_cap, ctypes.byref(self.layout))
            for sl in self.slots:
                sl["ws"] = None
            for sl in self.slots:
                sl["ws"] = torch.empty(self.layout.total, dtype=torch.uint8, device=eng.device)

    def buffer(self, k):
        """numpy view [cap_frames, N, 3] of slot k's pinned input buffer (decode straight into it)."""
        return self.slots[k]["in_np"]

    def submit(self, k, n_frames, src=None):
        """Queue the analysis of the first ``n_frames`` frames of slot k (H2D, kernels, D2H); returns at once.
        ``src``: a pinned [>= n_frames, N, 3] float64 torch tensor to read instead of the slot's own buffer
        (it must stay untouched until :meth:`collect` returns)."""
        sl, eng = self.slots[k], self.eng
        h_in = sl["h_in"] if src is None else src
        n = int(n_frames)
        if not 0 < n <= self.cap:
            raise ValueError(f"n_frames {n} outside 1..{self.cap}")
        if sl["ws"] is None:
            raise RuntimeError("ensure_ray_capacity() has not been called")
        if sl["busy"]:
            raise RuntimeError(f"slot {k} submitted again before it was collected")
        with torch.cuda.device(eng.device):
            p = self.params
            b, stream, _, _ = eng._make_batch(sl["d_in"], self.off, self.types, self.vdw, self.mass, True, self.N,
                                              None, p["stages"], p["adjust"], p["increment"], p["increment2"],
                                              p["pore_opt"], self.ray_cap)
            b.n_items, b.total_atoms = n, n * self.N
            need = eng.lib.pwb_workspace_bytes(n, n * self.N, self.N, self.ray_cap, eng.win_cap)
            if need > sl["ws"].numel():     # (huge molecules: a short batch spreads each item over more CTAs)
                sl["ws"] = torch.empty(need, dtype=torch.uint8, device=eng.device)
            b.workspace, b.workspace_bytes = sl["ws"].data_ptr(), sl["ws"].numel()
            _lib.check(eng.lib.pwb_full_analysis_host_async(
                eng._ctx, ctypes.byref(b), h_in.data_ptr(), sl["h_rec"].data_ptr(), sl["d_in"].data_ptr(),
                sl["d_rec"].data_ptr(), n * self.N, stream))
            sl["done"].record(torch.cuda.current_stream(eng.device))
            sl["n"], sl["src"], sl["busy"] = n, src, True

    def collect(self, k):
        """Wait for slot k's batch; numpy view [n_frames, rec_stride] of its records (valid until the slot is
        submitted again)."""
        sl = self.slots[k]
        if not sl["busy"]:
            raise RuntimeError(f"collect on slot {k} without a submitted batch")
        sl["done"].synchronize()
        sl["busy"] = False
        return sl["rec_np"][:sl["n"]]


_default_engine = None


def default_engine():
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine()
    return _default_engine


# ---------------------------------------------------------------------- record -> reference types
def sphere_volume(sphere_radius):
    """utilities.py:429-431."""
    return float(4 / 3 * np.pi * sphere_radius ** 3)


def record_windows(rec, warn=True):
    """(diameters[W], centres[W,3]) | None from one record, with find_windows' semantics
    (utilities.py:1473, :1515-1553): None if no ray is clear, dropped windows removed (warning),
    negative diameters kept (warning)."""
    status = int(rec[_lib.REC_STATUS])
    if status & _lib.ST_NO_CLEAR_RAY:
        return None
    nw = int(rec[_lib.REC_NWIN])
    w = np.asarray(rec[_lib.REC_WIN0: _lib.REC_WIN0 + 4 * nw], dtype=np.float64).reshape(nw, 4)
    keep = ~np.isnan(w[:, 0])
    if warn:
        for _ in range(int((~keep).sum())):
            logger.warning("Warning. One of the analysed windows has returned as None. See manual.")
        for _ in range(int((w[keep, 0] < 0).sum())):
            logger.warning("Warning. One of the analysed windows has a vdW corrected diameter smaller than 0. "
                           "See manual.")
    w = w[keep]
    return w[:, 0].copy(), w[:, 1:4].copy()


def check_record(rec):
    status = int(rec[_lib.REC_STATUS])
    if status & _lib.ST_BAD_ITEM:
        raise ValueError("molecule rejected by the GPU path (too many atoms for shared memory or unknown type)")
    if status & _lib.ST_RAY_CAP:
        raise RuntimeError("ray capacity exceeded (internal sizing error)")
    if status & _lib.ST_WIN_CAP:
        warnings.warn("more windows than win_cap; extra clusters were not analysed", RuntimeWarning, stacklevel=2)
    return status


def record_to_properties(rec, full=True):
    """Rebuild ``Molecule.properties`` (molecular.py:131-351) from one result record."""
    status = check_record(rec)
    p = {"no_of_atoms": int(rec[_lib.REC_NATOMS]),
         "centre_of_mass": np.array(rec[_lib.REC_COM:_lib.REC_COM + 3], dtype=np.float64),
         "maximum_diameter": {"diameter": float(rec[_lib.REC_MAXD]), "atom_1": int(rec[_lib.REC_MAXD_A1]),
                              "atom_2": int(rec[_lib.REC_MAXD_A2])},
         "average_diameter": float(rec[_lib.REC_AVGD]),
         "pore_diameter": {"diameter": float(rec[_lib.REC_PORE_D]), "atom": int(rec[_lib.REC_PORE_ATOM])},
         "pore_volume": sphere_volume(float(rec[_lib.REC_PORE_D]) / 2)}
    if status & _lib.ST_OPT_BOUNDS:
        # scipy's message for the inverted box the reference builds at utilities.py:415-420
        raise ValueError("An upper bound is less than the corresponding lower bound.")
    p["pore_diameter_opt"] = {"diameter": float(rec[_lib.REC_OPT_D]), "atom_1": int(rec[_lib.REC_OPT_ATOM]),
                              "centre_of_mass": np.array(rec[_lib.REC_OPT_C:_lib.REC_OPT_C + 3], dtype=np.float64)}
    p["pore_volume_opt"] = sphere_volume(float(rec[_lib.REC_OPT_D]) / 2)
    w = record_windows(rec)
    p["windows"] = ({"diameters": None, "centre_of_mass": None} if w is None
                    else {"diameters": w[0], "centre_of_mass": w[1]})
    return p
