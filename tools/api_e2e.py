"""End-to-end through the public Python API: DLPOLY(HISTORY).analysis() incl. text decode and dict building."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pywindow_b200 as pw
from pywindow_b200 import synthetic
G = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
F = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
seeds = np.stack([np.load(os.path.join(G, f"md_frame{i:02d}.npz"))["coords"] for i in range(20)])
el = np.load(os.path.join(G, "md_frame00.npz"))["elements"]
X = synthetic.noisy_frames(seeds, 0, F)
d = tempfile.mkdtemp()
path = os.path.join(d, "HISTORY")
t0 = time.perf_counter(); synthetic.write_history(path, [e + "1" for e in el], X); t1 = time.perf_counter()
print(f"wrote {F} frames, {os.path.getsize(path)/1e6:.0f} MB in {t1-t0:.1f} s")
t0 = time.perf_counter(); traj = pw.DLPOLY(path); t1 = time.perf_counter()
print(f"index: {t1-t0:.2f} s")
traj.analysis(frames=[0, 1], forcefield="DLF")          # warm up (engine, context)
traj.analysis_output = {}
t0 = time.perf_counter(); traj.analysis(forcefield="DLF"); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"analysis(): {t1-t0:.2f} s -> {F/(t1-t0):.0f} frames/s; 4 windows on {sum(len(v['0']['windows']['diameters'])==4 for v in traj.analysis_output.values())}/{F}")
import cProfile, pstats
traj.analysis_output = {}
pr = cProfile.Profile(); pr.enable(); traj.analysis(forcefield="DLF"); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
