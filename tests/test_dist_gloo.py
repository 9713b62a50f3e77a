"""world_size-2 (gloo, CPU) test of the multi-rank host logic: contiguous frame sharding in
Trajectory.analysis and the single end-of-run gather of result records.  The GPU engine is replaced
by a stub that fills records deterministically from the frame coordinates, so the test checks the
plumbing (which frames go where, record order after the gather), not the kernels."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden


class _StubResult:
    def __init__(self, rec):
        self._rec = rec

    def host_records(self):
        return self._rec


class _StubEngine:
    win_cap = 16
    rec_stride = 28 + 4 * 16

    def analyse_frames(self, coords, elements, **kw):
        from pywindow_b200 import _lib as L
        F, N = coords.shape[0], coords.shape[1]
        rec = np.zeros((F, self.rec_stride))
        rec[:, L.REC_NATOMS] = N
        rec[:, L.REC_COM:L.REC_COM + 3] = coords.mean(axis=1)
        rec[:, L.REC_MAXD] = coords.reshape(F, -1).max(axis=1)
        rec[:, L.REC_NWIN] = 1
        rec[:, L.REC_WIN0] = coords[:, 0, 0]
        return _StubResult(rec)

    def frame_pipeline(self, elements, cap_frames, ray_cap=0, slots=2, **params):
        return _StubPipeline(self, elements, cap_frames, slots)


class _StubPipeline:
    """Host-only stand-in for engine.FramePipeline: same buffer / submit / collect protocol."""

    vdw_max = 2.0

    def ensure_ray_capacity(self, ray_cap):
        assert ray_cap >= 64

    def __init__(self, eng, elements, cap, slots):
        self.eng, self.elements = eng, elements
        self.slots = [{"in": np.zeros((cap, len(elements), 3)), "rec": None} for _ in range(slots)]

    def buffer(self, k):
        return self.slots[k]["in"]

    def submit(self, k, n, src=None):
        sl = self.slots[k]
        assert sl["rec"] is None, "slot submitted again before it was collected"
        sl["rec"] = self.eng.analyse_frames(sl["in"][:n].copy(), self.elements).host_records()

    def collect(self, k):
        rec, self.slots[k]["rec"] = self.slots[k]["rec"], None
        assert rec is not None, "collect without submit"
        return rec


def _worker(rank, world, port, path, out):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pywindow_b200 import trajectory as T
    T.default_engine = lambda: _StubEngine()
    t = T.DLPOLY(path)
    t.batch_frames = 2                                  # 4 (3) frames per rank -> the two slots are reused
    t.analysis(frames="all", forcefield="DLF")
    keys = sorted(t.analysis_output)
    vals = np.array([t.analysis_output[k]["0"]["windows"]["diameters"][0] for k in keys])
    np.save(f"{out}_{rank}.npy", np.stack([np.array(keys, dtype=float), vals]))
    dist.destroy_process_group()


def test_sharded_analysis_equals_single_rank(tmp_path):
    import sys
    sys.path.insert(0, ROOT)
    from pywindow_b200 import synthetic
    from pywindow_b200 import trajectory as T
    g = [load_golden(f"md_frame{i:02d}") for i in range(4)]
    seeds = np.stack([x["coords"] for x in g])
    X = synthetic.noisy_frames(seeds, 0, 7)            # 7 frames over 2 ranks -> 4 + 3
    path = tmp_path / "HISTORY"
    synthetic.write_history(path, [e + "1" for e in g[0]["elements"]], X)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "res")
    mp.spawn(_worker, args=(2, port, str(path), out), nprocs=2, join=True)
    r0, r1 = np.load(out + "_0.npy"), np.load(out + "_1.npy")
    np.testing.assert_array_equal(r0, r1)               # every rank holds the full result
    assert list(r0[0]) == list(range(7))
    saved = T.default_engine
    T.default_engine = lambda: _StubEngine()
    try:
        t = T.DLPOLY(path)
        t.analysis(frames="all", forcefield="DLF")
    finally:
        T.default_engine = saved
    single = np.array([t.analysis_output[k]["0"]["windows"]["diameters"][0] for k in range(7)])
    np.testing.assert_array_equal(r0[1], single)


def _gather_worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pywindow_b200.distributed import gather_records
    n = 3 if rank == 0 else 1
    rec = np.arange(n * 5, dtype=np.float64).reshape(n, 5) + 100 * rank
    keys = [(10 * rank + i, i) for i in range(n)]
    R, K = gather_records(rec, keys)
    if rank == 0:
        np.save(out, R)
        assert K == [(0, 0), (1, 1), (2, 2), (10, 0)]
    dist.destroy_process_group()


def test_gather_records_ragged(tmp_path):
    """Direct test of gather_records with unequal shard sizes (3 + 1) and integer molecule keys."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "g.npy")
    mp.spawn(_gather_worker, args=(2, port, out), nprocs=2, join=True)
    R = np.load(out)
    assert R.shape == (4, 5) and R[3, 0] == 100.0
