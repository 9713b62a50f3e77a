"""Throughput of the periodic pipeline (config 3/4): PUDXES cell (1 344 atoms -> 8 x 168) + noise.
Under `ncu --metrics gpu__time_duration.sum` the launch list of the timed calls shows which kernels run between the
input copy and the record copy (profiles/r2_launches_periodic.csv)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pywindow_b200.engine import Engine
from pywindow_b200 import periodic, synthetic, _lib as L
from pywindow_b200.utilities import unit_cell_to_lattice_array
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "per_pudxes_cell.npz"))
lat = unit_cell_to_lattice_array(g["unit_cell"])
F = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
X = synthetic.noisy_frames(g["coords"][None], 0, F, decimals=3)
eng = Engine(0)
d_X = torch.as_tensor(X, device=eng.device)
for rep in range(reps):
    torch.cuda.synchronize(); t = time.perf_counter()
    db = periodic.discrete_molecules_batch(d_X, g["elements"], lattice=lat, rebuild=True, engine=eng)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    rec, keys, sizes = periodic.analyse_cells(d_X, g["elements"], lattice=lat, rebuild=True, engine=eng)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"F={F}: discrete_molecules {1e3*(t1-t):.1f} ms ({F/(t1-t):.0f} cells/s); full pipeline {1e3*(t2-t1):.1f} ms "
          f"({F/(t2-t1):.0f} cell-frames/s, {len(keys)/(t2-t1):.0f} molecule-frames/s); molecules/frame "
          f"{len(keys)/F:.2f}; 168 atoms: {(sizes==168).mean():.5f}; 4 windows: {(rec[:,L.REC_NWIN]==4).mean():.3f}")
