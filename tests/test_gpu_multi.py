"""Multi-GPU paths on real devices (needs >= 2 GPUs on the box; skipped otherwise): ray sharding of one huge
molecule over ranks (BASELINE config 5) and the sharded public trajectory call over NCCL."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ray_worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from pywindow_b200 import _lib as L
    from pywindow_b200.engine import Engine
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "big_saygor_adj113.npz")))
    eng = Engine(rank)
    adj = float(g["adjust"])
    res = eng.analyse_ray_sharded(g["elements"], g["coords"], adjust=adj, centre_override=g["fw_centre"])
    rec = res.host_records()[0]
    S = int(rec[L.REC_NCLEAR])
    np.savez(f"{out}_{rank}.npz", rec=rec, ck=res.clear_k.cpu().numpy()[0, :S], lab=res.labels.cpu().numpy()[0, :S],
             a2=res.clear_a2.cpu().numpy()[0, :S])
    if rank == 0:      # the same molecule on one GPU, for a bit-for-bit comparison
        one = eng.analyse_molecules([(g["elements"], g["coords"])], adjust=adj, centre_override=g["fw_centre"][None])
        np.save(f"{out}_single.npy", one.host_records()[0])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
def test_ray_sharding_over_two_gpus(tmp_path):
    """1e5 rays of SAYGOR split over 2 ranks: merged clear-ray list, DBSCAN labels and windows equal the reference
    goldens and the single-GPU record bit for bit, on every rank."""
    import torch.multiprocessing as mp
    from pywindow_b200 import _lib as L
    g = load_golden("big_saygor_adj113")
    out = str(tmp_path / "r")
    mp.spawn(_ray_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    single = np.load(out + "_single.npy")
    want = np.nonzero(g["ray_clear"])[0]
    for r in range(2):
        d = np.load(f"{out}_{r}.npz")
        assert (d["ck"] == want).all() and (d["lab"] == g["labels"]).all()
        np.testing.assert_allclose(d["a2"], g["ray_a2"], atol=1e-9)
        np.testing.assert_array_equal(d["rec"], single)
        nw = int(d["rec"][L.REC_NWIN])
        w = d["rec"][L.REC_WIN0:L.REC_WIN0 + 4 * nw].reshape(nw, 4)
        np.testing.assert_allclose(w[:, 0], g["windows_diam"], atol=1e-3)


def _traj_worker(rank, world, port, path, out):
    import sys
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["LOCAL_WORLD_SIZE"] = str(world)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import pywindow_b200 as pw
    t = pw.DLPOLY(path)
    t.batch_frames = 64
    t.analysis(forcefield="DLF")
    tab, keys = t.analysis_output.record_table()
    np.savez(f"{out}_{rank}.npz", tab=tab, frames=keys.frames, mols=keys.mols)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
def test_trajectory_analysis_sharded_over_nccl(tmp_path):
    """DLPOLY(path).analysis() on 2 ranks (frames sharded, records gathered over NCCL): every rank holds every
    frame, in order, with the records of a single-GPU analysis."""
    import torch.multiprocessing as mp
    from pywindow_b200 import synthetic
    from pywindow_b200.engine import Engine
    seeds = np.stack([load_golden(f"md_frame{i:02d}")["coords"] for i in range(20)])
    el = load_golden("md_frame00")["elements"]
    X = synthetic.noisy_frames(seeds, 0, 301)
    path = str(tmp_path / "HISTORY")
    synthetic.write_history_native(path, [e + "1" for e in el], X)
    out = str(tmp_path / "t")
    mp.spawn(_traj_worker, args=(2, _free_port(), path, out), nprocs=2, join=True)
    import pywindow_b200 as pw
    t = pw.DLPOLY(path)
    ref = Engine(0).analyse_frames(t._decode_coordinates(list(range(301))), el).host_records()
    for r in range(2):
        d = np.load(f"{out}_{r}.npz")
        assert (d["frames"] == np.arange(301)).all() and (d["mols"] == -1).all()
        np.testing.assert_array_equal(d["tab"], ref)
