"""Host build of the small optimisers used by the CUDA window / pore-centre kernels
(pywindow_b200/csrc/lbfgsb_small.cuh, neldermead2.cuh) against scipy -- the third-party code the
reference calls at utilities.py:422 (L-BFGS-B, n=3), :1301 (L-BFGS-B, n=1) and :1312 (brute+fmin).

The z-stage and Nelder-Mead restatements reproduce scipy bit for bit on every golden window; the
3-D pore-centre search is chaotic and is held to the reference's own noise floor."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
from scipy.optimize import fmin, minimize

from conftest import ROOT, golden_names, load_golden

CSRC = os.path.join(ROOT, "pywindow_b200", "csrc")
dp = ctypes.POINTER(ctypes.c_double)


def P(a):
    return a.ctypes.data_as(dp)


@pytest.fixture(scope="module")
def lib():
    so = os.path.join(CSRC, "libpwb_hostcheck.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so,
                           os.path.join(CSRC, "host_check.cpp")])
    L = ctypes.CDLL(so)
    L.pwbh_opt_z.argtypes = [dp, dp, ctypes.c_int, ctypes.c_double, dp]
    L.pwbh_fmin_xy.argtypes = [dp, dp, ctypes.c_int] + [ctypes.c_double] * 3 + [dp]
    L.pwbh_opt_pore.argtypes = [dp, dp, ctypes.c_int, dp, dp]
    return L


WIN = [n for n in golden_names(("cc3", "win_", "md_", "ex_")) ]


@pytest.mark.parametrize("name", WIN)
def test_z_stage_and_fmin_bitwise(lib, oracle, name):
    O = oracle
    g = load_golden(name)
    if "fw_centre" not in g or not bool(g["has_windows"]):
        pytest.skip("no windows")
    r, m, X = O.vdw_of(g["elements"]), O.masses_of(g["elements"]), g["coords"]
    det = {}
    O.find_windows(r, m, X, centre=g["fw_centre"], detail=det)
    for w in det["windows"]:
        if w["rescan"] is None:
            continue
        Y = np.ascontiguousarray(w["coords_rot"])
        nz = w["new_z"]
        zres = minimize(lambda z: O.pore_diameter(r, Y, np.array([0.0, 0.0, z[0]]))[0], x0=0.0,
                        bounds=[[-nz, None]])
        out = np.zeros(8)
        lib.pwbh_opt_z(P(Y), P(r), len(r), -nz, P(out))
        assert abs(out[0] - zres.x[0]) < 1e-12
        assert int(out[2]) == zres.nit
        d0, idx = w["d0"], w["grid_argmin"]
        gx = np.mgrid[slice(-d0 / 2, d0 / 2, 20j)]
        x0, y0, wz = gx[idx // 20], gx[idx % 20], zres.x[0]
        ref = fmin(lambda q: -O.pore_diameter(r, Y, np.array([q[0], q[1], wz]))[0], np.array([x0, y0]),
                   full_output=1, disp=0)
        o2 = np.zeros(8)
        lib.pwbh_fmin_xy(P(Y), P(r), len(r), wz, x0, y0, P(o2))
        assert np.abs(o2[:2] - ref[0]).max() < 1e-12
        assert int(o2[3]) == ref[3]


def test_pore_centre_search_noise_floor(lib, oracle):
    O = oracle
    diffs = []
    for name in WIN:
        g = load_golden(name)
        if "opt" not in g:
            continue
        r, m, X = O.vdw_of(g["elements"]), O.masses_of(g["elements"]), np.ascontiguousarray(g["coords"])
        com = O.center_of_mass(m, X)
        out = np.zeros(8)
        assert lib.pwbh_opt_pore(P(X), P(r), len(r), P(com), P(out)) == 0
        diffs.append(abs(-out[3] - g["opt"][0]))
    diffs = np.array(diffs)
    assert (diffs > 1e-3).mean() <= 0.4, diffs     # reference self-agreement under jitter: ~30 % off
    assert diffs.max() < 5e-2, diffs
