"""Time the stages of one library build on the bench workload and fingerprint its records.
usage: [PWB_LIBRARY=path/to/variant.so] python tools/stage_bench.py [frames] [reps]
Prints ms per stage per 10 000 frames (CUDA events, chunks=1) and a checksum of the free-running records, so
kernel variants (tools/variants.sh) can be compared for speed AND for unchanged results in one GPU call."""
import ctypes, hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pywindow_b200 import _lib as L
from pywindow_b200.engine import Engine, element_types, host_ray_capacity

F = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
el, seeds = bench.seed_frames()
cache = f"/tmp/pwb_synth_{F}.npy"
if os.path.exists(cache):
    X = np.load(cache)
else:
    X = bench.synth(seeds, 0, F); np.save(cache, X)
eng = Engine(0); eng.set_chunks(1)
dev = eng.device
N = len(el)
types, vdw, mass = element_types(el)
d_X = torch.as_tensor(X, device=dev).view(-1, 3)
off = torch.arange(0, (F + 1) * N, N, dtype=torch.int32, device=dev)
d_t, d_v, d_m = torch.as_tensor(types, device=dev), torch.as_tensor(vdw, device=dev), torch.as_tensor(mass, device=dev)
ray_cap = host_ray_capacity(X[:4096], float(vdw.max()))
lay = L.WsLayout(); eng.lib.pwb_workspace_layout(F, F * N, N, ray_cap, eng.win_cap, ctypes.byref(lay))
ws = torch.empty(lay.total, dtype=torch.uint8, device=dev)
rec = torch.empty((F, eng.rec_stride), dtype=torch.float64, device=dev)
sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
stages = [("prepare", L.STAGE_SCALARS | L.STAGE_AVGD), ("opt", L.STAGE_OPT), ("rays", L.STAGE_RAYS),
          ("cluster", L.STAGE_CLUSTER), ("windows", L.STAGE_WINDOWS)]
def batch(st):
    b, _, _, _ = eng._make_batch(d_X, off, d_t, d_v, d_m, True, N, None, st, 1.0, 1.0, 0.1, True, ray_cap)
    b.records, b.workspace, b.workspace_bytes = rec.data_ptr(), ws.data_ptr(), lay.total
    return b
bs = [batch(st) for _, st in stages]
tot = np.zeros(len(stages))
for r in range(reps + 1):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)]
    ev[0].record()
    for k, b in enumerate(bs):
        L.check(eng.lib.pwb_full_analysis_batch(eng._ctx, ctypes.byref(b), sp)); ev[k + 1].record()
    torch.cuda.synchronize()
    if r:
        tot += [ev[k].elapsed_time(ev[k + 1]) for k in range(len(stages))]
ms = tot / reps * 1e4 / F
h = rec.cpu().numpy()
fp = hashlib.sha1(np.ascontiguousarray(h).tobytes()).hexdigest()[:12]
print(os.environ.get("PWB_LIBRARY", "default"), "ms/10k:", " ".join(f"{n}={v:.3f}" for (n, _), v in zip(stages, ms)),
      f"sum={ms.sum():.3f}", "records", fp, "opt_d_sum %.9f" % h[:, L.REC_OPT_D].sum(), "nfev_mean %.1f" % h[:, L.REC_OPT_NFEV].mean(),
      "nfev p50/p99/p999/max %d/%d/%d/%d" % tuple(np.percentile(h[:, L.REC_OPT_NFEV], [50, 99, 99.9, 100])),
      "win4 %.4f" % (h[:, L.REC_NWIN] == 4).mean())
