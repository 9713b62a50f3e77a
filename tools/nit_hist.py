import sys, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
from pywindow_b200.engine import Engine
from pywindow_b200 import _lib as L
el, seeds = bench.seed_frames()
X = bench.synth(seeds, 0, 10000)
eng = Engine(0)
r = eng.analyse_frames(X, el, stages=L.STAGE_SCALARS|L.STAGE_OPT).host_records()
nf = r[:, L.REC_OPT_NFEV]; nit = r[:, L.REC_OPT_NIT]
print('nfev mean %.0f median %.0f p90 %.0f p99 %.0f max %.0f'%(nf.mean(), np.median(nf), np.percentile(nf,90), np.percentile(nf,99), nf.max()))
print('nit mean %.1f max %.0f'%(nit.mean(), nit.max()), 'stop codes', np.bincount(r[:, L.REC_OPT_STOP].astype(int)))
import time
for F in (148, 592, 2368, 10000):
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(5): eng.analyse_frames(X[:F], el, stages=L.STAGE_SCALARS|L.STAGE_OPT)
    torch.cuda.synchronize(); print(F, 'frames: %.3f ms'%((time.perf_counter()-t)/5*1e3))
