#!/usr/bin/env python
"""bench.py -- frames/s of the full window + pore analysis on synthetic CC3 MD trajectories.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], SURVEY.md 8d "config 2"): single-molecule CC3 MD frames,
168 atoms; frame f = seed[f mod 20] + N(0, 0.05 A) per coordinate with
numpy.random.default_rng(20261017 + f); the 20 seed frames are the frames of the reference's
examples/data/input/HISTORY_singlemol_short, shipped as tests/golden/md_frame*.npz.
One "step" = Molecule.full_analysis for every frame of one 10 000-frame batch per GPU.
Frames are independent, so ranks shard them with no data-path collective (weak scaling: every rank
analyses its own 10 000 frames per step); the records are gathered over NCCL once, after the last
step, inside the timed region.
"""
from __future__ import annotations

import argparse
import ctypes
import glob
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261017
SIGMA = 0.05
N_ROTATE = 4            # distinct input batches cycled through so inputs are not L2 resident


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
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU baseline (oracle)
def _cpu_one(args):
    el, X = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pw_oracle as O
    p = O.full_analysis(el, X)
    w = p["windows"]["diameters"]
    return 0 if w is None else len(w)


def cpu_baseline(elements, frames, cores):
    """Oracle port (oracle/pw_oracle.py) over `frames` on `cores` processes -> frames/s."""
    import multiprocessing as mp
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_one, [(elements, frames[0])] * cores)          # warm the workers (imports)
        t0 = time.perf_counter()
        pool.map(_cpu_one, [(elements, f) for f in frames], chunksize=1)
        dt = time.perf_counter() - t0
    return len(frames) / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    elements, seeds = seed_frames()
    cores = os.cpu_count() or 1
    per_step = max(16 * cores, 64)
    total_t, total_f = 0.0, 0
    fps = 0.0
    for step in range(args.warmup + args.steps):
        fr = synth(seeds, step * per_step, per_step)
        fps, dt = cpu_baseline(elements, fr, cores)
        if step >= args.warmup:
            total_t += dt
            total_f += per_step
    value = total_f / total_t
    sample = (f"{per_step} frames/step x {args.steps} steps of the same synthetic CC3 workload; oracle port "
              "(batched numpy + the reference's own scipy optimiser calls), one process per core; the unmodified "
              "reference measured 0.39 frames/s/core in the build container (BASELINE.md)")
    _emit(json.dumps({
        "impl": "reference", "metric": "frames/s full_analysis (CC3 traj)", "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "CC3 single-molecule MD trajectory, 168 atoms, full_analysis per frame (configs[1])",
                   "frames_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from pywindow_b200 import _lib as L
    from pywindow_b200.engine import Engine, element_types

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    eng = Engine(local)
    elements, seeds = seed_frames()
    N = len(elements)
    F = args.frames
    types, vdw, mass = element_types(elements)

    # this rank's frames: [rank*N_ROTATE*F, (rank+1)*N_ROTATE*F)
    host_batches = []
    for r in range(N_ROTATE):
        h = torch.from_numpy(synth(seeds, (rank * N_ROTATE + r) * F, F)).pin_memory()
        host_batches.append(h)
    dev_batches = [h.to(dev) for h in host_batches]
    off = torch.arange(0, (F + 1) * N, N, dtype=torch.int32, device=dev)
    d_types = torch.as_tensor(types, device=dev)
    d_vdw, d_mass = torch.as_tensor(vdw, device=dev), torch.as_tensor(mass, device=dev)

    # size the ray lists once (synchronising helper), then build one reusable batch descriptor
    first = eng.analyse(dev_batches[0].view(-1, 3), off, d_types, d_vdw, d_mass, shared_types=True, max_atoms=N)
    ray_cap = first.ray_cap + 64
    rec0 = first.host_records()
    layout = L.WsLayout()
    eng.lib.pwb_workspace_layout(F, F * N, ray_cap, eng.win_cap, ctypes.byref(layout))
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
    events = [[torch.cuda.Event(enable_timing=True) for _ in range(n_ev)] for _ in range(args.steps)]

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

    # ---- per-kernel durations: the same K steps again, stage by stage on one stream (chunks = 1),
    # CUDA events between the stage calls
    eng.set_chunks(1)
    step_staged(0, events[0])
    barrier()
    for i in range(args.steps):
        step_staged(args.warmup + i, events[i])
    barrier()
    stage_ms = np.array([[ev[k].elapsed_time(ev[k + 1]) for k in range(len(stage_list))] for ev in events]).mean(axis=0)
    eng.set_chunks(args.chunks)

    # ---- end-to-end: host (pinned) coords -> C ABI host entry -> host records, every step; one field of every
    # step's records is summed on the host before the next step is submitted.  (Keeping two batches in flight
    # with pwb_full_analysis_host_async was measured 6 % SLOWER: the interleaved kernels of two batches lose more
    # than the hidden copy latency gains, see DESIGN.md.)
    state = {}
    for i in range(max(1, args.warmup)):
        eng.analyse_frames_host(host_batches[i % N_ROTATE], elements, ray_cap=ray_cap, state=state)
    barrier()
    t0 = time.perf_counter()
    e2e_checksum = 0.0
    for i in range(args.steps):
        h_rec = eng.analyse_frames_host(host_batches[(args.warmup + i) % N_ROTATE], elements, ray_cap=ray_cap, state=state)
        e2e_checksum += float(h_rec[:, L.REC_PORE_D].sum())
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * F * args.steps / float(t.item())
    h2d = F * N * 3 * 8
    d2h = F * eng.rec_stride * 8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rank 0: roofline, CPU baseline, JSON line
    rec_h = rec.cpu().numpy()
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
    tj = os.path.join(ROOT, "profiles", "traffic.json")      # dram bytes per launch from the committed ncu capture
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
                       "frac": value / world * flop_frame / 1e12 / fp32.value},
        "hbm": {"achieved": value / world * bytes_frame / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": value / world * bytes_frame / 1e9 / hbm_peak,
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback"},
        "stages": stages,
        "stage_ms_note": "stage-by-stage re-run of the same steps on one stream (chunks=1); the timed region "
                         "pipelines %d sub-batches over internal streams, so its step time is below the sum" % args.chunks,
        "arithmetic": "f64 kernels; FLOP counted by the SURVEY 8(d) per-frame work formula",
    }
    cores = os.cpu_count() or 1
    n_cpu = 0 if args.skip_cpu else min(max(128 * cores, 256), 4096)      # ~10-15 s of all-core CPU work
    cpu_fps, cpu_dt = cpu_baseline(elements, synth(seeds, 0, n_cpu), cores) if n_cpu else (None, 0.0)
    nwin = rec_h[:, L.REC_NWIN]
    line = {
        "metric": "frames/s full_analysis (CC3 traj)", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "CC3 single-molecule MD trajectory, 168 atoms, full_analysis per frame (configs[1])",
                   "frames_per_step_per_gpu": F, "noise_sigma_A": SIGMA, "ray_cap": ray_cap, "chunks": args.chunks,
                   "l2": f"{N_ROTATE} distinct input batches cycled ({N_ROTATE * F * N * 24 / 1e6:.0f} MB) + "
                         f"{layout.total / 1e6:.0f} MB workspace rewritten every step (> 126 MB L2)",
                   "frames_with_4_windows": float((nwin == 4).mean())},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": {"value": cpu_fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{n_cpu} frames of the same synthetic workload in {cpu_dt:.1f} s; oracle port "
                                   "(batched numpy + scipy optimisers), one process per core"},
    }
    _emit(json.dumps(line))
    if world > 1:
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
    ap.add_argument("--frames", type=int, default=10000, help="frames per step per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-cpu", action="store_true", help="omit the cpu_baseline leg (profiling runs)")
    ap.add_argument("--chunks", type=int, default=4, help="sub-batches pipelined over internal streams per step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
