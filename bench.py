#!/usr/bin/env python
"""bench.py -- frames/s of the full window + pore analysis on synthetic CC3 MD trajectories.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], SURVEY.md 8d "config 2"): single-molecule CC3 MD frames,
168 atoms; frame f = seed[f mod 20] + N(0, 0.05 A) per coordinate with
numpy.random.default_rng(20261017 + f); the 20 seed frames are the frames of the reference's
examples/data/input/HISTORY_singlemol_short, shipped as tests/golden/md_frame*.npz.
One "step" = Molecule.full_analysis for every frame of one 100 000-frame batch per GPU (north_star's
trajectory length).  Frames are independent, so ranks shard them with no data-path collective (weak
scaling: every rank analyses its own batch per step); the records are gathered over NCCL once, after
the last step, inside the timed region.

The JSON line carries, beside the contract's keys:
  e2e        the same metric from pinned HOST buffers through the engine's pipelined host entry
             (engine.FramePipeline -> pwb_full_analysis_host_async: the entry Trajectory.analysis uses)
  api        frames/s of the public drop-in call pywindow_b200.DLPOLY(path).analysis() on a synthetic
             100 000-frame HISTORY file written in the run (file -> analysis_output; strong scaling over ranks)
  secondary  BASELINE configs 1, 3, 4, 5: CUDA-event / wall times, a parity bit against the reference-generated
             goldens, and the CPU port beside each
  roofline, cpu_baseline, clocks, gpu_launches
"""
from __future__ import annotations

import argparse
import ctypes
import glob
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261017
SIGMA = 0.05
N_ROTATE = 2            # distinct input batches cycled through (each one alone is larger than L2)
METRIC = "frames/s full_analysis (CC3 traj)"
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_STUB = os.path.join(ROOT, "oracle", "refstub")


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


def seed_frames():
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "md_frame*.npz")))
    g0 = np.load(files[0])
    return g0["elements"], np.stack([np.load(f)["coords"] for f in files])


def synth(seeds, first, count):
    out = np.empty((count,) + seeds.shape[1:], dtype=np.float64)
    for i in range(count):
        f = first + i
        out[i] = seeds[f % len(seeds)] + np.random.default_rng(SEED + f).normal(0.0, SIGMA, seeds.shape[1:])
    return out


def workload_config(frames):
    """The `config` object: identical in both arms (the reference arm times a bounded sample of it)."""
    return {"workload": "CC3 single-molecule MD trajectory, 168 atoms, full_analysis per frame (configs[1]) at "
                        "north_star's trajectory length",
            "frames_per_step_per_gpu": int(frames), "atoms": 168, "noise_sigma_A": SIGMA,
            "l2": f"every step reads a {frames * 168 * 24 / 1e6:.0f} MB input batch (two batches alternate) and "
                  "rewrites a multi-GB workspace: far beyond the 126 MB L2"}


def scratch_dir():
    for d in ("/dev/shm", tempfile.gettempdir()):
        try:
            st = os.statvfs(d)
            if os.access(d, os.W_OK) and st.f_bavail * st.f_frsize > 6e9:
                return d
        except OSError:
            pass
    return tempfile.gettempdir()


# ----------------------------------------------------------------------------- algorithmic work
def flop_model(rec, n_atoms, increment=1.0, increment2=0.1):
    """SURVEY.md 8(d) work formula evaluated with each frame's own P, S, C, W: mean algorithmic FLOP
    per frame, split by the stage (kernel) that does the work -> dict."""
    from pywindow_b200 import _lib as L
    N = n_atoms
    D = rec[:, L.REC_MAXD]
    R = D / 2
    Pw, S, W = rec[:, L.REC_NRAYS], rec[:, L.REC_NSURV], rec[:, L.REC_NWIN]
    C = np.floor(R / increment)
    C2 = np.floor(R / increment2)
    Pa = np.floor(np.log10(4 * np.pi * D * D) * 250)
    c_p, c_d, c_r, E_p, E_z, E_xy = 12.0, 11.0, 10.0, 100.0, 30.0, 70.0
    nclear = rec[:, L.REC_NCLEAR]
    parts = {
        "prepare+avg_diameter": c_p * N * N / 2 + c_d * N + c_r * N * Pa,
        "opt_pore": c_d * N * E_p + 0 * D,
        "rays": c_r * N * Pw + c_d * N * S * (C + 1),
        "cluster": 8 * nclear ** 2,
        "windows": W * c_d * N * (C2 + 1 + E_z + 400 + E_xy + 2),
    }
    out = {k: float(np.mean(v)) for k, v in parts.items()}
    out["total"] = float(sum(out.values()))
    return out


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.idx, self.rows, self.stop_flag, self.proc = gpu_index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in self.rows if len(r) > 8 and r[3].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None}


# ----------------------------------------------------------------------------- CPU baselines
def _cpu_one(args):
    el, X = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pw_oracle as O
    p = O.full_analysis(el, X)
    w = p["windows"]["diameters"]
    return 0 if w is None else len(w)


def cpu_port(elements, frames, cores):
    """Oracle port (oracle/pw_oracle.py) over `frames` on `cores` processes -> (frames/s, seconds)."""
    import multiprocessing as mp
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_one, [(elements, frames[0])] * cores)          # warm the workers (imports)
        t0 = time.perf_counter()
        pool.map(_cpu_one, [(elements, f) for f in frames], chunksize=1)
        dt = time.perf_counter() - t0
    return len(frames) / dt, dt


def cpu_port_fresh(n_frames, cores):
    """cpu_port in a fresh interpreter with single-threaded BLAS: a process that has already used numpy's BLAS
    (or CUDA) forks workers that oversubscribe the cores 16-fold."""
    code = ("import sys, json; sys.path.insert(0, %r); import bench; el, seeds = bench.seed_frames(); "
            "fps, dt = bench.cpu_port(el, bench.synth(seeds, 0, %d), %d); print('PORTJSON' + json.dumps([fps, dt]))"
            % (ROOT, n_frames, cores))
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=1800)
    for line in out.stdout.splitlines():
        if line.startswith("PORTJSON"):
            return tuple(json.loads(line[8:]))
    raise RuntimeError("port run failed: " + out.stderr[-500:])


def reference_available():
    return os.path.isdir(os.path.join(REF_DIR, "pywindow"))


_REF_CHILD = r"""
import json, os, sys, time
sys.path[:0] = [{stub!r}, {ref!r}]
os.environ.setdefault("OMP_NUM_THREADS", "1")
import pywindow as pw
path, per_step, steps, warmup, cores = {path!r}, {per_step}, {steps}, {warmup}, {cores}
t0 = time.perf_counter()
traj = pw.DLPOLY(path)
t_map = time.perf_counter() - t0
times, nwin = [], []
for s in range(warmup + steps):
    frames = list(range(s * per_step, (s + 1) * per_step))
    t0 = time.perf_counter()
    traj.analysis(frames=frames, ncpus=cores, ncpus_analysis=1, forcefield="DLF")
    dt = time.perf_counter() - t0
    if s >= warmup:
        times.append(dt)
    nwin += [len(traj.analysis_output[f]["0"]["windows"]["diameters"]) for f in frames]
print("REFJSON" + json.dumps({{"seconds": times, "map_s": t_map, "nwin4": sum(n == 4 for n in nwin) / len(nwin),
                              "version": getattr(pw, "__version__", "?"), "file": pw.__file__}}))
"""


def reference_run(elements, seeds, per_step, steps, warmup, cores):
    """The UNMODIFIED reference (pip-installed into oracle/_ref by __graft_entry__.build()): its own
    DLPOLY(path).analysis(frames=..., ncpus=cores) (trajectory.py:350-586) on a HISTORY file of the same synthetic
    workload, in a fresh interpreter (its multiprocessing.Pool forks; no CUDA context in that process)."""
    from pywindow_b200 import synthetic
    n = per_step * (steps + warmup)
    d = tempfile.mkdtemp(prefix="pwb_ref_")
    path = os.path.join(d, "HISTORY")
    synthetic.write_history_native(path, [str(e) + "1" for e in elements], synth(seeds, 0, n), compact=False)
    code = _REF_CHILD.format(stub=REF_STUB, ref=REF_DIR, path=path, per_step=per_step, steps=steps, warmup=warmup, cores=cores)
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    try:
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=3000)
    finally:
        pass
    for line in out.stdout.splitlines():
        if line.startswith("REFJSON"):
            r = json.loads(line[7:])
            os.remove(path)
            os.rmdir(d)
            return r
    raise RuntimeError("reference run failed: " + out.stderr[-800:])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    elements, seeds = seed_frames()
    cores = os.cpu_count() or 1
    base = {"impl": "reference", "metric": METRIC, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args.frames)}
    if reference_available():
        per_step = 2 * cores
        r = reference_run(elements, seeds, per_step, args.steps, args.warmup, cores)
        total_t = float(sum(r["seconds"]))
        value = per_step * args.steps / total_t
        kind = "reference"
        sample = (f"{per_step} frames/step x {args.steps} steps of the same synthetic CC3 workload, read from a HISTORY "
                  f"file by the unmodified reference: pywindow.DLPOLY(path).analysis(frames=..., ncpus={cores}, "
                  f"ncpus_analysis=1, forcefield='DLF') (trajectory.py:350-586, multiprocessing.Pool); 4 windows on "
                  f"{100 * r['nwin4']:.0f} % of frames; OMP_NUM_THREADS=1")
    else:
        per_step = max(16 * cores, 64)
        total_t = 0.0
        for step in range(args.warmup + args.steps):
            _, dt = cpu_port(elements, synth(seeds, step * per_step, per_step), cores)   # fresh process already
            if step >= args.warmup:
                total_t += dt
        value = per_step * args.steps / total_t
        kind = "port"
        sample = (f"{per_step} frames/step x {args.steps} steps; oracle port (batched numpy + the reference's own scipy "
                  "optimiser calls), one process per core -- oracle/_ref (the installed reference) is missing")
    base.update({"value": value, "ms_per_step": 1e3 * total_t / args.steps,
                 "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
                 "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    _emit(json.dumps(base))


# ----------------------------------------------------------------------------- secondary configs
def _cuda_ms(torch, fn, reps):
    """Mean CUDA-event time of fn() on the current stream (after one warm call)."""
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, 1e3 * (time.perf_counter() - t0) / reps


def secondary_rank0(torch, eng, skip_cpu):
    """BASELINE configs 1, 3, 5 on one GPU: time, parity bit against the reference-generated goldens, CPU port."""
    import pywindow_b200 as pw
    from pywindow_b200 import _lib as L, periodic
    from pywindow_b200.engine import record_windows
    from pywindow_b200.utilities import unit_cell_to_lattice_array
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    out = {}

    # ---- config 1: CC3 cage, Molecule.full_analysis() through the drop-in API
    g = golden("cc3")
    mol = pw.MolecularSystem.load_system({"elements": g["elements"], "coordinates": g["coords"]}, "cc3").system_to_molecule()
    dev_ms, wall_ms = _cuda_ms(torch, lambda: mol.full_analysis(), 20)
    p = mol.full_analysis()
    w = p["windows"]["diameters"]
    ok = (w is not None and len(w) == len(g["windows_diam"]) and
          float(np.abs(np.sort(w) - np.sort(g["windows_diam"])).max()) < 1e-3 and
          abs(p["maximum_diameter"]["diameter"] - g["maxdim"][2]) < 1e-9 and
          abs(p["pore_diameter"]["diameter"] - g["pore"][0]) < 1e-9 and
          abs(p["average_diameter"] - float(g["avg_diameter"])) < 1e-9 and
          abs(p["pore_diameter_opt"]["diameter"] - g["opt"][0]) < 1e-3)
    c1 = {"what": "Molecule.full_analysis() of the CC3 cage (tests/test_validate_cc3.py system), 168 atoms, 797 rays",
          "latency_ms": wall_ms, "cuda_event_ms": dev_ms, "parity_vs_reference_golden": bool(ok),
          "windows": None if w is None else [float(x) for x in w]}
    if not skip_cpu:
        import pw_oracle as O
        t0 = time.perf_counter()
        O.full_analysis(g["elements"], g["coords"])
        c1["cpu_port_ms"] = 1e3 * (time.perf_counter() - t0)
        c1["cpu_reference_ms_survey"] = 2200.0
    out["config1_cc3_single"] = c1

    # ---- config 3: periodic PUDXES cell: discrete_molecules (rebuild) + full_analysis of its 8 molecules
    g = golden("per_pudxes_noise1")
    lat = unit_cell_to_lattice_array(g["unit_cell"])
    X1 = np.ascontiguousarray(g["coords"][None])
    dm_ms, _ = _cuda_ms(torch, lambda: periodic.discrete_molecules_batch(X1, g["elements"], lattice=lat, rebuild=True, engine=eng), 10)
    full_ms, full_wall = _cuda_ms(torch, lambda: periodic.analyse_cells(X1, g["elements"], lattice=lat, rebuild=True, engine=eng), 10)
    res = periodic.discrete_molecules_batch(X1, g["elements"], lattice=lat, rebuild=True, engine=eng)
    mols = res.molecules(0, g["elements"])
    co = np.concatenate([m["coordinates"] for m in mols])
    rec, keys, sizes = periodic.analyse_cells(X1, g["elements"], lattice=lat, rebuild=True, engine=eng)
    ok = (len(mols) == int(g["rebuild_count"]) and [len(m["elements"]) for m in mols] == list(g["rebuild_sizes"]) and
          float(np.abs(co - g["rebuild_coords"]).max()) < 1e-7 and bool((rec[:, L.REC_NWIN] == 4).all()))
    # batch throughput of the same stage: 2 048 noisy copies of the cell
    from pywindow_b200 import synthetic
    Fc = 2048
    Xc = synthetic.noisy_frames(golden("per_pudxes_cell")["coords"][None], 0, Fc, decimals=3)
    d_Xc = torch.as_tensor(Xc, device=eng.device)
    dmb_ms, _ = _cuda_ms(torch, lambda: periodic.discrete_molecules_batch(d_Xc, g["elements"], lattice=lat, rebuild=True, engine=eng), 3)
    pipe_ms, pipe_wall = _cuda_ms(torch, lambda: periodic.analyse_cells(d_Xc, g["elements"], lattice=lat, rebuild=True, engine=eng), 3)
    c3 = {"what": "PUDXES cell (1 344 atoms, cubic 24.8 A): discrete_molecules(rebuild) -> 8 x 168 atoms, then full_analysis per molecule",
          "single_cell": {"discrete_molecules_ms": dm_ms, "full_pipeline_ms": full_ms, "full_pipeline_wall_ms": full_wall},
          "batch": {"cells": Fc, "discrete_molecules_cells_per_s": Fc / (dmb_ms * 1e-3),
                    "full_pipeline_cell_frames_per_s": Fc / (pipe_wall * 1e-3),
                    "molecule_frames_per_s": 8 * Fc / (pipe_wall * 1e-3)},
          "parity_vs_reference_golden": bool(ok)}
    if not skip_cpu:
        import pw_periodic_oracle as PO
        t0 = time.perf_counter()
        PO.discrete_molecules(g["elements"], g["coords"], lattice=lat, rebuild=True)
        c3["cpu_port_discrete_molecules_ms"] = 1e3 * (time.perf_counter() - t0)
        c3["cpu_reference_cell_frame_s_survey"] = 22.0
    out["config3_periodic_cell"] = c3
    del d_Xc

    # ---- config 5: dense ray sampling (adjust = 113 -> ~1e5 rays) on the SAYGOR metal-organic cage
    g = golden("big_saygor_adj113")
    adj = float(g["adjust"])
    f5 = lambda: eng.analyse_molecules([(g["elements"], g["coords"])], adjust=adj).host_records()  # noqa: E731
    dev_ms, wall_ms = _cuda_ms(torch, f5, 5)
    forced = eng.analyse_molecules([(g["elements"], g["coords"])], centre_override=g["fw_centre"][None], adjust=adj)
    r = forced.host_records()[0]
    S = int(r[L.REC_NCLEAR])
    want = np.nonzero(g["ray_clear"])[0]
    w = record_windows(r, warn=False)
    ok = (int(r[L.REC_NRAYS]) == g["ray_clear"].shape[0] and S == len(want) and
          bool((forced.clear_k.cpu().numpy()[0, :S] == want).all()) and
          bool((forced.labels.cpu().numpy()[0, :S] == g["labels"]).all()) and
          abs(r[L.REC_AVGD] - float(g["avg_diameter"])) < 1e-9 and w is not None and
          len(w[0]) == len(g["windows_diam"]) and float(np.abs(w[0] - g["windows_diam"]).max()) < 1e-3)
    c5 = {"what": "SAYGOR (448 atoms) find_windows + find_average_diameter + pore analysis with adjust=113",
          "rays": int(r[L.REC_NRAYS]), "clear_rays": S, "windows": int(r[L.REC_NWIN]), "latency_ms": wall_ms,
          "cuda_event_ms": dev_ms, "parity_vs_reference_golden": bool(ok),
          "parity_note": "teacher-forced with the reference's pore centre: ray count, clear-ray list, DBSCAN labels exact; "
                         "average diameter 1e-9; window diameters 1e-3 A"}
    if not skip_cpu:
        import pw_oracle as O
        gm = golden("big_mibqar_adj8")
        rv, m = O.vdw_of(gm["elements"]), O.masses_of(gm["elements"])
        t0 = time.perf_counter()
        O.find_windows(rv, m, gm["coords"], adjust=float(gm["adjust"]), centre=gm["fw_centre"])
        c5["cpu_port_ms_bounded_sample"] = 1e3 * (time.perf_counter() - t0)
        c5["cpu_port_sample"] = "oracle find_windows on MIBQAR with adjust=8 (7 124 rays); the 1e5-ray case takes minutes on the CPU"
        f8 = lambda: eng.analyse_molecules([(gm["elements"], gm["coords"])], adjust=float(gm["adjust"])).host_records()  # noqa: E731
        c5["gpu_ms_same_sample"] = _cuda_ms(torch, f8, 5)[1]
    out["config5_dense_rays"] = c5
    return out


def api_leg(torch, dist, world, rank, path, n_frames, kwargs, warm_frames=64):
    """frames/s of the public call DLPOLY(path).analysis(**kwargs): file -> analysis_output.  Under torch.distributed
    the frame list is sharded by Trajectory.analysis itself and the records are gathered over NCCL at the end;
    the time is the max over ranks."""
    import pywindow_b200 as pw
    warm = pw.DLPOLY(path)
    warm.analysis(frames=list(range(min(warm_frames * world, n_frames))), **kwargs)   # contexts, pinned pools, NCCL
    del warm
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    traj = pw.DLPOLY(path)
    t1 = time.perf_counter()
    traj.analysis(**kwargs)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    t = torch.tensor([t2 - t0, t1 - t0, t2 - t1], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total, t_map, t_an = (float(v) for v in t.tolist())
    return traj, total, t_map, t_an


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from pywindow_b200 import _lib as L
    from pywindow_b200 import synthetic
    from pywindow_b200.engine import Engine, element_types, host_ray_capacity
    import pywindow_b200.engine as E

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    eng = Engine(local)
    E._default_engine = eng                      # the public API (DLPOLY.analysis, Molecule) uses this rank's engine
    elements, seeds = seed_frames()
    N = len(elements)
    F = args.frames
    types, vdw, mass = element_types(elements)

    # this rank's frames: [rank*N_ROTATE*F, (rank+1)*N_ROTATE*F)
    host_batches = []
    for r in range(N_ROTATE):
        host_batches.append(torch.from_numpy(synth(seeds, (rank * N_ROTATE + r) * F, F)).pin_memory())
    dev_batches = [h.to(dev) for h in host_batches]
    off = torch.arange(0, (F + 1) * N, N, dtype=torch.int32, device=dev)
    d_types = torch.as_tensor(types, device=dev)
    d_vdw, d_mass = torch.as_tensor(vdw, device=dev), torch.as_tensor(mass, device=dev)

    # ray-list capacity from the frames' bounding boxes (host side, as Trajectory.analysis does)
    ray_cap = host_ray_capacity(host_batches[0].numpy()[:4096], float(vdw.max()))
    layout = L.WsLayout()
    eng.lib.pwb_workspace_layout(F, F * N, N, ray_cap, eng.win_cap, ctypes.byref(layout))
    ws = torch.empty(layout.total, dtype=torch.uint8, device=dev)
    rec = torch.empty((F, eng.rec_stride), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev)
    sp = ctypes.c_void_p(stream.cuda_stream)

    def make(coords, stages):
        b, _, _, _ = eng._make_batch(coords.view(-1, 3), off, d_types, d_vdw, d_mass, True, N, None, stages,
                                     1.0, 1.0, 0.1, True, ray_cap)
        b.records, b.workspace, b.workspace_bytes = rec.data_ptr(), ws.data_ptr(), layout.total
        return b

    stage_list = [("prepare+avg_diameter", L.STAGE_SCALARS | L.STAGE_AVGD), ("opt_pore", L.STAGE_OPT),
                  ("rays", L.STAGE_RAYS), ("cluster", L.STAGE_CLUSTER), ("windows", L.STAGE_WINDOWS)]
    full = [make(c, L.STAGE_ALL) for c in dev_batches]
    staged = [[make(c, st) for _, st in stage_list] for c in dev_batches]
    n_ev = len(stage_list) + 1
    n_staged = min(args.steps, 5)
    events = [[torch.cuda.Event(enable_timing=True) for _ in range(n_ev)] for _ in range(n_staged)]

    def step(i):
        """One step = Molecule.full_analysis for every frame of the batch: one C-ABI call; the
        library pipelines `chunks` sub-batches over its internal streams."""
        L.check(eng.lib.pwb_full_analysis_batch(eng._ctx, ctypes.byref(full[i % N_ROTATE]), sp))

    def step_staged(i, ev):
        ev[0].record(stream)
        for k, b in enumerate(staged[i % N_ROTATE]):
            L.check(eng.lib.pwb_full_analysis_batch(eng._ctx, ctypes.byref(b), sp))
            ev[k + 1].record(stream)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    eng.set_chunks(args.chunks)
    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = eng.launches
    gathered = None
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record(stream)
    for i in range(args.steps):
        step(args.warmup + i)
    if world > 1:   # results gathered once, at the end, over NCCL / NVLink
        gathered = torch.empty((world * F, eng.rec_stride), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(gathered, rec)
    t_end.record(stream)
    barrier()
    ms = t_start.elapsed_time(t_end)
    launches = eng.launches - launches0
    clocks = sampler.finish() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * F * args.steps / (ms_max * 1e-3)
    del gathered

    # ---- per-kernel durations: a few of the same steps again, stage by stage on one stream (chunks = 1),
    # CUDA events between the stage calls
    eng.set_chunks(1)
    step_staged(0, events[0])
    barrier()
    for i in range(n_staged):
        step_staged(args.warmup + i, events[i])
    barrier()
    stage_ms = np.array([[ev[k].elapsed_time(ev[k + 1]) for k in range(len(stage_list))] for ev in events]).mean(axis=0)
    eng.set_chunks(args.chunks)
    rec_h = rec.cpu().numpy() if rank == 0 else None
    del full, staged, ws, rec
    torch.cuda.empty_cache()

    # ---- end to end: pinned HOST coordinates -> engine.FramePipeline (pwb_full_analysis_host_async, the entry
    # Trajectory.analysis uses) -> pinned host records, every step; two slots, so the H2D copy of step i+1 runs
    # under the kernels of step i; one field of every step's records is summed on the host (the D2H read).
    pipe = eng.frame_pipeline(elements, F, ray_cap, slots=2)
    e2e_checksum = 0.0
    for i in range(max(2, args.warmup)):
        pipe.submit(i % 2, F, src=host_batches[i % N_ROTATE])
        pipe.collect(i % 2)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        pipe.submit(i % 2, F, src=host_batches[(args.warmup + i) % N_ROTATE])
        if i:
            e2e_checksum += float(pipe.collect((i - 1) % 2)[:, L.REC_PORE_D].sum())
    e2e_checksum += float(pipe.collect((args.steps - 1) % 2)[:, L.REC_PORE_D].sum())
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * F * args.steps / float(t.item())
    h2d = F * N * 3 * 8
    d2h = F * eng.rec_stride * 8
    del pipe, dev_batches
    torch.cuda.empty_cache()

    # ---- public API: DLPOLY(path).analysis() on a 100 000-frame HISTORY written in the run (strong scaling)
    api = None
    cfg4 = None
    if not args.skip_api:
        d = scratch_dir()
        tag = os.environ.get("MASTER_PORT", str(os.getpid()))
        path = os.path.join(d, f"pwb_bench_{tag}.HISTORY")
        n_api = args.api_frames
        if rank == 0:
            X = host_batches[0].numpy()[:n_api] if n_api <= F else synth(seeds, 0, n_api)
            size = synthetic.write_history_native(path, [str(e) + "1" for e in elements], X, compact=True)
        barrier()
        try:
            # under torch.distributed the warm-up covers the whole file, so that NCCL's buffers and the gather's
            # staging have the size of the timed call (their first use costs more than the call itself)
            traj, total, t_map, t_an = api_leg(torch, dist, world, rank, path, n_api, {"forcefield": "DLF"},
                                               warm_frames=64 if world == 1 else n_api)
            if rank == 0:
                tab, keys = traj.analysis_output.record_table()
                n4 = float((tab[:, L.REC_NWIN] == 4).mean())
                # the decoder alone, all host threads of this rank, into a preallocated buffer
                from pywindow_b200.trajectory import _host_threads
                nd = min(n_api, 20000)
                buf = np.empty((nd, N, 3))
                t0 = time.perf_counter()
                traj._decode_into(list(range(nd)), buf, _host_threads())
                dec = nd / (time.perf_counter() - t0)
                mm, view = traj._mapped()                       # the host's plain streaming rate over the same bytes
                nbytes = int(traj.trajectory_map.pairs[nd - 1, 1])
                t0 = time.perf_counter()
                eng.lib.pwb_stream_probe(view.ctypes.data, nbytes, _host_threads())
                stream_gbs = nbytes / (time.perf_counter() - t0) / 1e9
                dec_gbs = dec * (size / n_api) / 1e9
                api = {"value": n_api / total, "unit": "frames/s", "frames": n_api, "scaling": "strong", "n_gpus": world,
                       "call": "pywindow_b200.DLPOLY(path).analysis(forcefield='DLF')  (file -> analysis_output)",
                       "seconds": {"total": total, "map_file": t_map, "analysis": t_an},
                       "analysis_only_frames_per_s": n_api / t_an,
                       "file": {"bytes": int(size), "format": "DL_POLY HISTORY, keytrj 0, imcon 0, %12.4E records "
                                "(the layout of the reference's HISTORY_singlemol_short)"},
                       "frames_analysed": len(keys), "frames_with_4_windows": n4,
                       "host_decode_frames_per_s": dec, "host_threads": _host_threads(),
                       "host_decode_GBps": dec_gbs, "host_stream_GBps": stream_gbs,
                       "pools": "pinned / device buffers warm from an earlier analysis() in this process (engine pool); "
                                "N > 1: the earlier call covered the whole file (NCCL buffers of the gather's size)",
                       "limit": "the GPU at the reader's batch size (16 384 frames per batch, two sub-batches each: "
                                "about 1.4e6 frames/s from pinned buffers against 1.6e6 at 100 000 frames per batch) "
                                "plus the head (map, first decode) and tail (last batch, record read-back) of a 0.1 s "
                                "job; the decoder (host_decode_frames_per_s on all host threads of the rank; "
                                "host_stream_GBps = the same threads merely counting line feeds) runs under the "
                                "kernels; ranks of one node share the host cores"}
            del traj
        finally:
            barrier()
            if rank == 0 and os.path.exists(path):
                os.remove(path)
        # ---- config 4: periodic trajectory through the same public call (modular, rebuild)
        gp = golden("per_pudxes_cell")
        from pywindow_b200.utilities import unit_cell_to_lattice_array
        import pywindow_b200.trajectory as pw_traj
        lat = unit_cell_to_lattice_array(gp["unit_cell"])
        n_cells = args.cell_frames
        path4 = os.path.join(d, f"pwb_bench_{tag}_cells.HISTORY")
        if rank == 0:
            Xc = synthetic.noisy_frames(gp["coords"][None], 0, n_cells, decimals=3)
            size4 = synthetic.write_history_native(path4, [str(e) + "1" for e in gp["elements"]], Xc, lattice=lat, compact=True)
            del Xc
        barrier()
        try:
            kw4 = {"modular": True, "rebuild": True, "forcefield": "DLF"}
            # warm-up = one full batch of cells per rank, so that the pooled device / pinned buffers have their final
            # size (as for the api leg above: the timed call is a second analysis() of the process)
            traj, total, t_map, t_an = api_leg(torch, dist, world, rank, path4, n_cells, kw4,
                                               warm_frames=pw_traj.Trajectory.modular_batch_cells)
            if rank == 0:
                tab, keys = traj.analysis_output.record_table()
                cfg4 = {"what": "periodic PUDXES cell trajectory (1 344 atoms/cell, 8 cages): DLPOLY(path).analysis("
                                "modular=True, rebuild=True, forcefield='DLF'), file -> analysis_output",
                        "cell_frames": n_cells, "n_gpus": world, "scaling": "strong",
                        "pools": "warm (second analysis() of the process, as for api)",
                        "cell_frames_per_s": n_cells / total, "molecule_frames_per_s": len(keys) / total,
                        "seconds": {"total": total, "map_file": t_map, "analysis": t_an}, "file_bytes": int(size4),
                        "parity": {"molecules_per_cell": len(keys) / n_cells,
                                   "all_168_atoms": bool((tab[:, L.REC_NATOMS] == 168).all()),
                                   "natoms_hist": {str(int(a)): int(c) for a, c in
                                                   zip(*np.unique(tab[:, L.REC_NATOMS], return_counts=True))},
                                   "frames_with_4_windows": float((tab[:, L.REC_NWIN] == 4).mean())},
                        "cpu_reference_cell_frames_per_s_survey": 0.045}
            del traj
        finally:
            barrier()
            if rank == 0 and os.path.exists(path4):
                os.remove(path4)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- rank 0: secondary configs, roofline, CPU baseline, JSON line
    secondary = None
    if not args.skip_secondary:
        secondary = secondary_rank0(torch, eng, args.skip_cpu)
        if cfg4 is not None:
            secondary["config4_periodic_trajectory"] = cfg4
    flops = flop_model(rec_h, N)
    flop_frame = flops["total"]
    fp32 = ctypes.c_double(0)
    fp64 = ctypes.c_double(0)
    L.check(eng.lib.pwb_fma_peak(eng._ctx, 0, sp, ctypes.byref(fp32)))
    L.check(eng.lib.pwb_fma_peak(eng._ctx, 1, sp, ctypes.byref(fp64)))
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = {}
    tj = os.path.join(ROOT, "profiles", "traffic.json")      # dram bytes per frame from the committed ncu capture
    if os.path.exists(tj):
        traffic = json.load(open(tj))
    names = [n for n, _ in stage_list]
    kernel_of = {"prepare+avg_diameter": "k_prepare", "opt_pore": "k_opt_pore", "rays": "k_rays",
                 "cluster": "k_cluster", "windows": "k_windows"}
    stages = {}
    for n, ms_ in zip(names, stage_ms):
        tf = flops[n] * F / (ms_ * 1e-3) / 1e12
        tr = traffic.get(kernel_of[n])
        stages[n] = {"kernel": kernel_of[n], "ms": float(ms_), "share": float(ms_ / stage_ms.sum()),
                     "flop_per_frame": flops[n], "achieved": tf, "frac": tf / fp32.value,
                     "traffic": (tr["bytes_per_frame"] * F if tr else None)}
    dom = names[int(np.argmax(stage_ms))]
    bytes_frame = N * 3 * 8 + eng.rec_stride * 8
    roofline = {
        "bound": "fp32-fma (compute; neither hbm nor tensor: K=3 contractions, 4.7 KB/frame)",
        "kernel": kernel_of[dom], "achieved": stages[dom]["achieved"], "peak": fp32.value, "unit": "TFLOP/s",
        "frac": stages[dom]["frac"], "traffic": stages[dom]["traffic"],
        "definition": "dominant kernel by time; achieved = SURVEY 8(d) algorithmic FLOP of that stage x frames / its "
                      "CUDA-event duration; opt_pore / windows are serial optimisers (latency bound), rays / prepare "
                      "are FP-pipe bound",
        "peak_source": "pwb_fma_peak FMA micro-benchmark in this run (MEASURED_PEAKS.json has no FP32 figure); "
                       "nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4",
        "fp64_peak_measured": fp64.value,
        "whole_step": {"flop_per_frame": flop_frame, "achieved": value / world * flop_frame / 1e12,
                       "frac": value / world * flop_frame / 1e12 / fp32.value,
                       "frac_of_fp64_peak": value / world * flop_frame / 1e12 / fp64.value},
        "hbm": {"achieved": value / world * bytes_frame / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": value / world * bytes_frame / 1e9 / hbm_peak,
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback"},
        "stages": stages,
        "stage_ms_note": "stage-by-stage re-run of %d of the same steps on one stream (chunks=1); the timed region "
                         "pipelines %d sub-batches over internal streams, so its step time is below the sum" % (n_staged, args.chunks),
        "arithmetic": "f64 kernels; FLOP counted by the SURVEY 8(d) per-frame work formula",
    }
    cores = os.cpu_count() or 1
    cpu = {"value": None, "unit": "frames/s", "cores": cores, "kind": "port", "sample": "skipped"}
    if not args.skip_cpu:
        n_cpu = min(max(96 * cores, 256), 3072)                 # ~10 s of all-core CPU work for the port
        port_fps, port_dt = cpu_port_fresh(n_cpu, cores)
        cpu = {"value": port_fps, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": f"{n_cpu} frames of the same synthetic workload in {port_dt:.1f} s; oracle port "
                         "(batched numpy + scipy optimisers), one process per core"}
        if reference_available():
            try:
                per = 2 * cores
                r = reference_run(elements, seeds, per, 2, 0, cores)
                ref_fps = 2 * per / float(sum(r["seconds"]))
                cpu = {"value": ref_fps, "unit": "frames/s", "cores": cores, "kind": "reference",
                       "sample": f"{2 * per} frames of the same synthetic workload from a HISTORY file through the "
                                 f"unmodified reference pywindow.DLPOLY(path).analysis(ncpus={cores}) "
                                 f"(trajectory.py:350-586) in {sum(r['seconds']):.1f} s; OMP_NUM_THREADS=1",
                       "port": {"value": port_fps, "sample_frames": n_cpu, "port_over_reference": port_fps / ref_fps}}
            except Exception as exc:     # noqa: BLE001 -- the baseline is reported, never fatal
                cpu["reference_error"] = str(exc)[-300:]
    nwin = rec_h[:, L.REC_NWIN]
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(F),
        "run": {"ray_cap": ray_cap, "chunks": args.chunks, "workspace_bytes": int(layout.total),
                "frames_with_4_windows": float((nwin == 4).mean()), "e2e_checksum": e2e_checksum},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "entry": "engine.FramePipeline -> pwb_full_analysis_host_async (pinned host coords in, pinned host "
                         "records out, two slots)"},
        "api": api,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "secondary": secondary,
    }
    _emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner from C
# when the first communicator is created), so file descriptor 1 is pointed at stderr for the whole run and the
# line goes to the saved descriptor.
_REAL_STDOUT = None


def _capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(text):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (text + "\n").encode())


def main():
    _capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=100000, help="frames per step per GPU")
    ap.add_argument("--api-frames", type=int, default=100000, help="frames of the HISTORY file of the api leg")
    ap.add_argument("--cell-frames", type=int, default=4096, help="cell-frames of the periodic trajectory (config 4)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-cpu", action="store_true", help="omit the cpu_baseline leg (profiling runs)")
    ap.add_argument("--skip-api", action="store_true", help="omit the public-API legs (profiling runs)")
    ap.add_argument("--skip-secondary", action="store_true", help="omit the secondary configs (profiling runs)")
    ap.add_argument("--chunks", type=int, default=2, help="sub-batches pipelined over internal streams per step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
