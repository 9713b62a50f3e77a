"""TEST INFRASTRUCTURE ONLY -- import the *reference* pywindow (read-only at /root/reference).

Only usable in the build container (the GPU box has no /root/reference).  Used by
``oracle/make_golden.py`` to generate the committed fixtures under ``tests/golden/`` and by
the optional ``-m "not gpu"`` tests that cross-check the oracle restatement live.

The reference needs ``rdkit`` only for ``rdkit.Chem.inchi.logger`` (utilities.py:11); a stub
package lives in ``oracle/refstub``.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys

REF_ROOT = "/root/reference"
_STUB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refstub")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "src", "pywindow"))


def load_reference():
    """Return ``(pywindow, pywindow._internal.utilities)`` of the reference."""
    if not reference_available():
        raise RuntimeError("reference not present (expected only in the build container)")
    for p in (_STUB, os.path.join(REF_ROOT, "src")):
        if p not in sys.path:
            sys.path.insert(0, p)
    pw = importlib.import_module("pywindow")
    util = importlib.import_module("pywindow._internal.utilities")
    return pw, util


def load_reference_test_module(name: str):
    """Import ``/root/reference/tests/<name>.py`` (for its literal input arrays)."""
    load_reference()
    path = os.path.join(REF_ROOT, "tests", name + ".py")
    spec = importlib.util.spec_from_file_location("_reftest_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class Recorder:
    """Wraps the reference's stage functions so one unmodified ``find_windows`` run yields
    every intermediate needed for stage-wise (teacher-forced) parity tests.

    Wrapped names (all looked up as module globals by the reference at call time):
    ``opt_pore_diameter`` (utilities.py:400), ``max_dim`` (:355), ``vector_preanalysis`` (:1132),
    ``vector_analysis`` (:1100), ``DBSCAN`` (:1478), ``window_analysis`` (:1191),
    ``minimize`` (:1301, :422), ``brute`` (:1312).
    """

    def __init__(self, util):
        self.u = util
        self.saved = {}
        self.reset()

    def reset(self):
        self.rays = []          # per ray: None | 8-vector
        self.ray_points = []
        self.db = None          # (eps, dataset, labels)
        self.windows = []       # per cluster dict
        self.opt = []           # opt_pore_diameter returns
        self.maxdim = []
        self._cur = None
        self.shifted = None

    def __enter__(self):
        u = self.u
        names = ["opt_pore_diameter", "max_dim", "vector_preanalysis", "vector_analysis",
                 "DBSCAN", "window_analysis", "minimize", "brute"]
        self.saved = {n: getattr(u, n) for n in names}
        s = self.saved
        rec = self

        def opt_pore_diameter(*a, **k):
            r = s["opt_pore_diameter"](*a, **k)
            rec.opt.append(r)
            return r

        def max_dim(el, co):
            r = s["max_dim"](el, co)
            rec.maxdim.append(r)
            rec.shifted = co.copy()
            return r

        def vector_preanalysis(v, co, vdw, increment=1.0):
            r = s["vector_preanalysis"](v, co, vdw, increment=increment)
            rec.rays.append(None if r is None else r.copy())
            rec.ray_points.append(v.copy())
            return r

        def vector_analysis(v, co, vdw, increment=1.0):
            r = s["vector_analysis"](v, co, vdw, increment=increment)
            if rec._cur is not None and "rescan" not in rec._cur:
                rec._cur["rescan_vector"] = v.copy()
                rec._cur["rescan"] = None if r is None else r.copy()
            return r

        class DBSCAN:  # noqa: N801
            def __init__(self, eps):
                self.eps = eps
                self._o = s["DBSCAN"](eps=eps)

            def fit(self, data):
                o = self._o.fit(data)
                rec.db = (self.eps, data.copy(), o.labels_.copy())
                return o

        def minimize(fun, x0=None, args=(), bounds=None, **k):
            r = s["minimize"](fun, x0=x0, args=args, bounds=bounds, **k)
            if rec._cur is not None:
                rec._cur["z_x0"] = float(x0)
                rec._cur["z_lb"] = bounds[0][0]
                rec._cur["z_opt"] = float(r.x[0])
                rec._cur["z_fun"] = float(r.fun)
                rec._cur["z_nfev"] = int(r.nfev)
                rec._cur["z_nit"] = int(r.nit)
                rec._cur["coords_rot"] = args[3].copy()
            return r

        def brute(fun, ranges, args=(), full_output=False, finish=None, **k):
            r = s["brute"](fun, ranges, args=args, full_output=True, finish=finish, **k)
            if rec._cur is not None:
                jout = r[3]
                idx = int(jout.ravel().argmin())
                rec._cur["xy_bounds"] = float(ranges[0][1])
                rec._cur["grid_argmin"] = idx
                rec._cur["grid_min"] = float(jout.ravel()[idx])
                rec._cur["xy_opt"] = r[0].copy()
                rec._cur["xy_fun"] = float(r[1])
            return r if full_output else r[0]

        def window_analysis(window, el, co, vdw, **k):
            rec._cur = {"cluster_rows": window.copy()}
            r = s["window_analysis"](window, el, co, vdw, **k)
            cur = rec._cur
            rec._cur = None
            cur["result"] = None if r is None else (float(r[0]), r[1].copy())
            rec.windows.append(cur)
            return r

        for n, f in dict(opt_pore_diameter=opt_pore_diameter, max_dim=max_dim,
                         vector_preanalysis=vector_preanalysis, vector_analysis=vector_analysis,
                         DBSCAN=DBSCAN, window_analysis=window_analysis, minimize=minimize,
                         brute=brute).items():
            setattr(u, n, f)
        return self

    def __exit__(self, *exc):
        for n, f in self.saved.items():
            setattr(self.u, n, f)
        return False
