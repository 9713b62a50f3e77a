"""cProfile of the host side of analyse_cells (config 4 pipeline)."""
import os, sys, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pywindow_b200.engine import Engine
from pywindow_b200 import periodic, synthetic
from pywindow_b200.utilities import unit_cell_to_lattice_array
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "per_pudxes_cell.npz"))
lat = unit_cell_to_lattice_array(g["unit_cell"])
F = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
X = synthetic.noisy_frames(g["coords"][None], 0, F, decimals=3)
eng = Engine(0)
for _ in range(2):
    periodic.analyse_cells(X, g["elements"], lattice=lat, rebuild=True, engine=eng)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
periodic.analyse_cells(X, g["elements"], lattice=lat, rebuild=True, engine=eng)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
