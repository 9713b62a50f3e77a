"""How far does the pore-centre optimiser travel from its start (the centre of mass)?  Sizes the
register-resident near list of k_opt_pore."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pywindow_b200.engine import Engine
from pywindow_b200 import _lib as L
el, seeds = bench.seed_frames()
X = bench.synth(seeds, 0, 10000)
eng = Engine(0)
r = eng.analyse_frames(X, el, stages=L.STAGE_SCALARS | L.STAGE_OPT).host_records()
com = r[:, L.REC_COM:L.REC_COM + 3]; c = r[:, L.REC_OPT_C:L.REC_OPT_C + 3]
d = np.linalg.norm(c - com, axis=1)
print("travel |c_opt - com|: mean %.3f median %.3f p90 %.3f p99 %.3f max %.3f" % (d.mean(), np.median(d), *np.percentile(d, [90, 99]), d.max()))
print("pore radius: mean %.3f min %.3f max %.3f" % ((r[:, L.REC_PORE_D] / 2).mean(), (r[:, L.REC_PORE_D] / 2).min(), (r[:, L.REC_PORE_D] / 2).max()))
# how many atoms have clearance <= a(com) + 2 rho at the start point?
from pywindow_b200.engine import element_types
types, vdw, mass = element_types(el)
rad = np.asarray(vdw)[np.asarray(types)]
Xn = np.asarray(X)
for rho in (0.25, 0.5, 0.75, 1.0, 1.5, 2.0):
    cnt = []
    for f in range(0, 10000, 50):
        cl = np.linalg.norm(Xn[f] - com[f], axis=1) - rad
        cnt.append(int((cl <= cl.min() + 2 * rho).sum()))
    cnt = np.array(cnt)
    print("rho %.2f: atoms kept mean %.1f p90 %d max %d; frames with travel <= rho: %.1f %%" % (
        rho, cnt.mean(), np.percentile(cnt, 90), cnt.max(), 100 * (d <= rho).mean()))
