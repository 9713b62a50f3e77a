"""Single-molecule latencies: config 1 (CC3 full_analysis) and config 5 (SAYGOR, adjust=113, ~1e5 rays)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pywindow_b200 as pw
from pywindow_b200.engine import default_engine
from pywindow_b200 import _lib as L
G = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
def t(fn, n=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
g = np.load(os.path.join(G, "cc3.npz"))
mol = pw.MolecularSystem.load_system({"elements": g["elements"], "coordinates": g["coords"]}, "cc3").system_to_molecule()
print("config 1  CC3 Molecule.full_analysis(): %.2f ms" % t(lambda: mol.full_analysis()))
g = np.load(os.path.join(G, "big_saygor_adj113.npz"))
eng = default_engine()
f = lambda: eng.analyse_molecules([(g["elements"], g["coords"])], adjust=113.0).host_records()
ms = t(f, 3)
r = f()[0]
print("config 5  SAYGOR adjust=113: %.1f ms  (rays %d, clear %d, windows %d, avg diameter rays ~%d)" % (
    ms, r[L.REC_NRAYS], r[L.REC_NCLEAR], r[L.REC_NWIN], int(np.log10(4 * np.pi * r[L.REC_MAXD] ** 2) * 250 * 113)))
