"""Robustness run: hollow random 'cages' of very different sizes through the whole pipeline (shared-memory
sizing, warps-per-CTA fallbacks, ray caps); prints status flags, window counts and timings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pywindow_b200.engine import Engine
from pywindow_b200 import _lib as L
rng = np.random.default_rng(11)
eng = Engine(0)
for n in (12, 60, 168, 500, 1200, 2000, 2800, 3200, 5000, 10000, 20000):   # > ~2 900: atom tiles in global scratch
    R = 2.2 * np.sqrt(n / 12.0)                     # shell radius so that atoms stay ~3 A apart
    mols = []
    for k in range(6 if n <= 5000 else 2):
        v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1)[:, None]
        X = v * (R + rng.normal(0, 0.3, size=(n, 1)))
        el = rng.choice(["C", "H", "N", "O", "S", "Cl", "Br"], size=n, p=[.4, .35, .08, .08, .03, .03, .03])
        mols.append((list(el), X))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = eng.analyse_molecules(mols)
    rec = out.host_records(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"N={n:5d}: {1e3*(t1-t0):8.2f} ms  status {sorted(set(int(s) for s in rec[:, L.REC_STATUS]))} "
          f"rays {int(rec[0, L.REC_NRAYS])} windows {[int(w) for w in rec[:, L.REC_NWIN]]} "
          f"maxdim {rec[0, L.REC_MAXD]:.1f} pore {rec[0, L.REC_PORE_D]:.2f} opt {rec[0, L.REC_OPT_D]:.2f} avg {rec[0, L.REC_AVGD]:.2f}")
    assert np.isfinite(rec[:, [L.REC_MAXD, L.REC_PORE_D, L.REC_AVGD]]).all()
print("ok")
