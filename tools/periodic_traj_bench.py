"""Config 4 through the public call, several times in one process: DLPOLY(path).analysis(modular=True, rebuild=True)
on a synthetic periodic PUDXES trajectory; prints wall time per call and the time inside the cell decoder and
`periodic.analyse_cells` (first call = cold pools)."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pywindow_b200 as pw
from pywindow_b200 import periodic, synthetic, trajectory
from pywindow_b200.utilities import unit_cell_to_lattice_array
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "per_pudxes_cell.npz"))
lat = unit_cell_to_lattice_array(g["unit_cell"])
F = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
path = os.path.join(d, "pwb_cells_tool.HISTORY")
X = synthetic.noisy_frames(g["coords"][None], 0, F, decimals=3)
synthetic.write_history_native(path, [str(e) + "1" for e in g["elements"]], X, lattice=lat, compact=True)
acc = {}
def timed(mod, name):
    fn = getattr(mod, name)
    def w(*a, **k):
        t = time.perf_counter(); r = fn(*a, **k); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t; return r
    setattr(mod, name, w)
timed(periodic, "analyse_cells"); timed(pw.DLPOLY, "_decode_cells")
try:
    for rep in range(reps):
        acc.clear()
        t = time.perf_counter(); tr = pw.DLPOLY(path); t1 = time.perf_counter()
        tr.analysis(modular=True, rebuild=True, forcefield="DLF"); torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"rep {rep}: map {1e3*(t1-t):.1f} ms, analysis {1e3*(t2-t1):.1f} ms ({F/(t2-t):.0f} cell-frames/s); "
              + ", ".join(f"{k} {1e3*v:.1f} ms" for k, v in acc.items()), flush=True)
finally:
    os.remove(path)
