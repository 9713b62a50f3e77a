"""TEST INFRASTRUCTURE ONLY -- measure the reference's OWN free-running noise floor and commit it as a fixture.

Run in the build container:  ``python oracle/make_noise_floor.py``   (-> tests/golden/noise_floor.npz)

``opt_pore_diameter`` (utilities.py:400-426) is scipy L-BFGS-B with forward differences on a non-smooth max-min
objective: where it stops is chaotic.  Everything downstream of the pore centre when it is not
forced -- the ray origin of ``find_windows`` and hence window diameters / centres -- inherits that.  Parity of a
free-running re-implementation can therefore only be stated against the reference's agreement WITH ITSELF under a
perturbation far below any physical meaning.  This script runs the unmodified reference twice on every frame:

    base      X                                 (20 MD frames of HISTORY_singlemol_short + the first 200 frames of the
    jittered  X + U(-1e-12, 1e-12) per coordinate    synthetic bench workload, seeds 20261017 + f)

and stores, per frame and run: pore_diameter_opt (diameter, centre), the window count, window diameters and centres
(``find_windows`` free-running, cluster order).  Tests compare |GPU - base| and |oracle - base| with |jittered - base|.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

OUT = os.path.join(ROOT, "tests", "golden", "noise_floor.npz")
N_SYNTH = 200
WCAP = 8
JITTER = 1e-12


def frames():
    import glob
    from pywindow_b200 import synthetic
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "md_frame*.npz")))
    el = np.load(files[0])["elements"]
    seeds = np.stack([np.load(f)["coords"] for f in files])
    X = np.concatenate([seeds, synthetic.noisy_frames(seeds, 0, N_SYNTH)])
    return el, X


def one(args):
    el, X = args
    from refharness import load_reference
    _, util = load_reference()
    od, oa, oc = util.opt_pore_diameter(el, X)
    w = util.find_windows(el, X)
    wd = np.full(WCAP, np.nan)
    wc = np.full((WCAP, 3), np.nan)
    n = -1
    if w is not None:
        n = len(w[0])
        wd[:n] = w[0][:WCAP]
        wc[:n] = np.asarray(w[1])[:WCAP]
    return od, oa, np.asarray(oc), n, wd, wc


def main():
    el, X = frames()
    rng = np.random.default_rng(20261017)
    Xj = X + rng.uniform(-JITTER, JITTER, X.shape)
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    with mp.get_context("fork").Pool(max(1, (os.cpu_count() or 2) - 1)) as pool:
        res = pool.map(one, [(el, x) for x in X] + [(el, x) for x in Xj], chunksize=1)
    F = X.shape[0]
    out = {"n_md": np.array(20), "n_synth": np.array(N_SYNTH), "jitter": np.array(JITTER), "elements": el.astype("U4"),
           "coords_checksum": np.array([X.sum(), np.abs(X).sum()]), "jitter_offsets": Xj - X}
    for tag, part in (("base", res[:F]), ("jit", res[F:])):
        out[tag + "_opt_d"] = np.array([r[0] for r in part])
        out[tag + "_opt_atom"] = np.array([r[1] for r in part])
        out[tag + "_opt_c"] = np.stack([r[2] for r in part])
        out[tag + "_nwin"] = np.array([r[3] for r in part])
        out[tag + "_win_d"] = np.stack([r[4] for r in part])
        out[tag + "_win_c"] = np.stack([r[5] for r in part])
    import scipy
    import sklearn
    out["versions"] = np.array([f"numpy {np.__version__}", f"scipy {scipy.__version__}", f"sklearn {sklearn.__version__}"])
    np.savez_compressed(OUT, **out)
    d = np.abs(out["jit_opt_d"] - out["base_opt_d"])
    wdiff = np.nanmax(np.abs(out["jit_win_d"] - out["base_win_d"]), axis=1)
    print(f"{F} frames; reference vs itself under +-{JITTER:g} jitter: pore_opt > 1e-3 A on {(d > 1e-3).mean():.3f} of "
          f"frames (worst {d.max():.3e}); window diameters > 1e-3 A on {(wdiff > 1e-3).mean():.3f} (worst {wdiff.max():.3e}); "
          f"window count equal on {(out['jit_nwin'] == out['base_nwin']).mean():.3f}")


if __name__ == "__main__":
    main()
