"""Where the time of the public call goes: DLPOLY(path).analysis() on a synthetic CC3 HISTORY file (default 100 000
frames), single process or under torchrun; wall-clock timers around the host-side steps (decode, submit, collect,
gather, output assembly).  usage: [torchrun ...] python tools/api_profile.py [frames] [reps]"""
import os, sys, time, glob, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
import pywindow_b200 as pw
from pywindow_b200 import synthetic, trajectory, distributed, engine
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
F = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
files = sorted(glob.glob(os.path.join(root, "md_frame*.npz")))
seeds = np.stack([np.load(f)["coords"] for f in files]); el = [str(e) for e in np.load(files[0])["elements"]]
d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
path = os.path.join(d, "pwb_api_profile.HISTORY")
if rank == 0:
    synthetic.write_history_native(path, [e + "1" for e in el], synthetic.noisy_frames(seeds, 0, F), compact=True)
if world > 1: dist.barrier()
acc = {}
if os.environ.get("PWB_BATCH_FRAMES"): trajectory.Trajectory.batch_frames = int(os.environ["PWB_BATCH_FRAMES"])
def timed(obj, name, label=None):
    fn = getattr(obj, name); label = label or name
    def w(*a, **k):
        t = time.perf_counter(); r = fn(*a, **k); acc[label] = acc.get(label, 0.0) + time.perf_counter() - t; return r
    setattr(obj, name, w)
timed(pw.DLPOLY, "_decode_into"); timed(pw.DLPOLY, "_elements_of"); timed(trajectory.Trajectory, "_redo_overflowed")
timed(trajectory.Trajectory, "_analyse_single"); timed(trajectory.AnalysisOutput, "_add_block")
_orig_submit = engine.FramePipeline.submit
_events = []
def _submit(self, k, n, src=None):
    t = time.perf_counter()
    _orig_submit(self, k, n, src)
    e = torch.cuda.Event(enable_timing=True); e.record(torch.cuda.current_stream()); _events.append((t, n, e))
engine.FramePipeline.submit = _submit
timed(engine.FramePipeline, "submit"); timed(engine.FramePipeline, "collect"); timed(engine.FramePipeline, "ensure_ray_capacity")
timed(engine, "host_ray_capacity")
import pywindow_b200.trajectory as T
if hasattr(T, "gather_records"): timed(T, "gather_records")
timed(distributed, "gather_records", "gather_records(d)")
try:
    for rep in range(reps):
        acc.clear()
        if world > 1: dist.barrier()
        t = time.perf_counter(); tr = pw.DLPOLY(path); t1 = time.perf_counter()
        tr.analysis(forcefield="DLF"); torch.cuda.synchronize(); t2 = time.perf_counter()
        if world > 1: dist.barrier()
        print(f"[rank {rank}] rep {rep}: map {1e3*(t1-t):.1f} ms, analysis {1e3*(t2-t1):.1f} ms ({F/(t2-t):.0f} frames/s); "
              + ", ".join(f"{k} {1e3*v:.1f}" for k, v in acc.items()), flush=True)
        if rank == 0 and rep == reps - 1 and _events:
            ev = _events[-(len(_events) // reps):] if len(_events) % reps == 0 else _events
            t0 = ev[0][0]
            print("   batches (host submit time ms, frames, GPU done relative to first done ms): " +
                  "; ".join(f"{1e3*(t - t0):.1f}/{n}/{ev[0][2].elapsed_time(e):.1f}" for t, n, e in ev), flush=True)
        del tr
finally:
    if world > 1: dist.barrier()
    if rank == 0: os.remove(path)
    if world > 1: dist.destroy_process_group()
