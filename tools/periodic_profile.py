"""Where the periodic pipeline (analyse_cells) spends its time: wall-clock per host step with synchronisation
between them (2 048 noisy PUDXES cells)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pywindow_b200.engine import Engine
from pywindow_b200 import periodic, synthetic, _lib as L
from pywindow_b200.utilities import unit_cell_to_lattice_array
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "per_pudxes_cell.npz"))
lat = unit_cell_to_lattice_array(g["unit_cell"])
F = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
X = synthetic.noisy_frames(g["coords"][None], 0, F, decimals=3)
eng = Engine(0)
el = g["elements"]
N = X.shape[1]
periodic.analyse_cells(X, el, lattice=lat, rebuild=True, engine=eng)
def sync(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(2):
    t0 = sync()
    d_coords, d_lat, shared = periodic._device_inputs(eng, X, lat)
    t1 = sync()
    out_cap = int(max(64, 2.0 * N))
    b = periodic._dm_buffers(eng, F, N, out_cap, el)
    stream = periodic._run_dm(eng, b, d_coords, d_lat, shared, F, N, out_cap, True, 0.4)
    t2 = sync()
    items, atoms, biggest, st = periodic._compact(eng, b, F, N, out_cap, stream)
    t3 = sync()
    out = eng.analyse(b["coords"][:atoms], b["item_off"][:items + 1], b["dense_types"][:atoms], b["a_vdw"], b["a_mass"],
                      shared_types=False, max_atoms=biggest)
    t4 = sync()
    rec = out.host_records()
    t5 = sync()
    fi = b["item_frame"][:items].cpu().numpy(); mi = b["item_mol"][:items].cpu().numpy()
    sizes = np.diff(b["item_off"][:items + 1].cpu().numpy())
    t6 = sync()
    print(f"F={F}: H2D {1e3*(t1-t0):.2f}  discrete_molecules {1e3*(t2-t1):.2f}  compact {1e3*(t3-t2):.2f}  analyse "
          f"{1e3*(t4-t3):.2f}  records D2H {1e3*(t5-t4):.2f}  keys D2H {1e3*(t6-t5):.2f}  total {1e3*(t6-t0):.2f} ms; items {items}")
