"""TEST INFRASTRUCTURE ONLY -- generate ``tests/golden/*.npz`` by running the *reference*.

Run in the build container:  ``python oracle/make_golden.py [--only NAME]``

Every fixture holds the inputs (elements, coordinates) of one molecule, the reference's
public results for it (``Molecule.full_analysis`` pieces, molecular.py:156-202) and the
per-stage intermediates of one unmodified ``find_windows`` run (utilities.py:1364-1553),
captured with ``oracle/refharness.Recorder`` so the CUDA path can be tested stage by stage
("teacher forced": SURVEY.md section 8(c)).  Library versions are stored in each file.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from refharness import REF_ROOT, Recorder, load_reference, load_reference_test_module  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
NAN8 = np.full(8, np.nan)


def writer_system(g):
    """The system both writers are given in the writer fixtures / tests."""
    keys = np.array([e + str(i % 7) for i, e in enumerate(g["elements"])])
    return {"elements": g["elements"].copy(), "atom_ids": keys, "coordinates": g["coords"] * 3.7 - 11.0,
            "unit_cell": np.array([24.8, 25.1, 26.0, 90.0, 95.5, 101.25])}


def versions():
    import scipy
    import sklearn
    return np.array([f"numpy {np.__version__}", f"scipy {scipy.__version__}",
                     f"sklearn {sklearn.__version__}"])


def analyse(pw, util, elements, coords, name, windows=True, scalars=True, avg=True, adjust=1):
    elements = np.asarray(elements)
    coords = np.asarray(coords, dtype=np.float64)
    out = {"elements": elements.astype("U4"), "coords": coords, "versions": versions()}
    if scalars:
        out["com"] = util.center_of_mass(elements, coords)
        i1, i2, d = util.max_dim(elements, coords)
        out["maxdim"] = np.array([i1, i2, d], dtype=np.float64)
        pd_, pa = util.pore_diameter(elements, coords)
        out["pore"] = np.array([pd_, pa], dtype=np.float64)
        try:
            od, oa, oc = util.opt_pore_diameter(elements, coords)
            out["opt"] = np.array([od, oa, *oc], dtype=np.float64)
        except ValueError as e:  # bounds inverted when pore_diameter(COM) < 0
            out["opt_error"] = np.array([str(e)])
    if avg:
        out["avg_diameter"] = np.array(util.find_average_diameter(elements, coords, adjust=adjust))
        out["adjust"] = np.array(float(adjust))
    if windows and "opt_error" not in out:
        with Recorder(util) as rec:
            res = util.find_windows(elements, coords, adjust=adjust)
        P = len(rec.rays)
        out["fw_centre"] = np.asarray(rec.opt[0][2], dtype=np.float64)
        out["fw_shifted"] = rec.shifted
        out["fw_maxdim"] = np.array(rec.maxdim[0], dtype=np.float64)
        out["ray_points"] = np.array(rec.ray_points).reshape(P, 3)
        clear = np.array([r is not None for r in rec.rays], dtype=bool)
        out["ray_clear"] = clear
        out["ray_records"] = np.array([r for r in rec.rays if r is not None]).reshape(-1, 8)
        if rec.db is not None:
            out["eps"] = np.array(rec.db[0])
            out["labels"] = rec.db[2].astype(np.int64)
        W = len(rec.windows)
        wk = {k: [] for k in ("rescan_vector", "rescan", "z_lb", "z_opt", "z_fun", "z_nfev", "z_nit",
                              "xy_bounds", "grid_argmin", "grid_min", "xy_opt", "xy_fun",
                              "diam", "com_local")}
        for w in rec.windows:
            wk["rescan_vector"].append(w["rescan_vector"])
            wk["rescan"].append(NAN8 if w["rescan"] is None else w["rescan"])
            for k in ("z_lb", "z_opt", "z_fun", "z_nfev", "z_nit", "xy_bounds", "grid_argmin",
                      "grid_min", "xy_fun"):
                wk[k].append(w.get(k, np.nan))
            wk["xy_opt"].append(w.get("xy_opt", np.full(2, np.nan)))
            if w["result"] is None:
                wk["diam"].append(np.nan)
                wk["com_local"].append(np.full(3, np.nan))
            else:
                wk["diam"].append(w["result"][0])
                wk["com_local"].append(w["result"][1])
        for k, v in wk.items():
            out["win_" + k] = np.array(v, dtype=np.float64).reshape(W, -1) if W else np.zeros((0, 1))
        if res is None:
            out["has_windows"] = np.array(False)
        else:
            out["has_windows"] = np.array(True)
            out["windows_diam"] = np.asarray(res[0], dtype=np.float64)
            out["windows_com"] = np.asarray(res[1], dtype=np.float64).reshape(-1, 3)
    if adjust != 1 and "ray_records" in out:          # dense ray sets: keep the fixtures small
        out["ray_a2"] = out["ray_records"][:, 1].copy()
        for k in ("ray_records", "ray_points", "fw_shifted", "win_rescan_vector"):
            out.pop(k, None)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    nw = out.get("windows_diam", np.zeros(0)).shape[0]
    print(f"{name}: N={len(elements)} windows={nw}", flush=True)


def periodic(pw, name, path, noise_seed=None):
    """discrete_molecules / rebuild fixtures (utilities.py:820-1085, molecular.py:672-824).
    With ``noise_seed`` the cell gets N(0, 0.05 A) noise rounded to 3 decimals (a PDB-precision MD
    frame): tie-free, so atom and molecule ORDER is pinned as well."""
    import copy
    out = {"versions": versions()}
    ms = pw.MolecularSystem.load_file(path)
    sysd = ms.system
    if noise_seed is not None:
        rng = np.random.default_rng(noise_seed)
        sysd["coordinates"] = np.round(sysd["coordinates"] + rng.normal(0, 0.05, sysd["coordinates"].shape), 3)
    base = copy.deepcopy(sysd)
    _load = pw.MolecularSystem.load_file
    pw.MolecularSystem.load_file = classmethod(lambda cls, p: cls.load_system(copy.deepcopy(base), "noisy")) \
        if noise_seed is not None else _load
    out["elements"] = np.asarray(sysd["elements"]).astype("U4")
    out["coords"] = np.asarray(sysd["coordinates"], dtype=np.float64)
    out["unit_cell"] = np.asarray(sysd["unit_cell"], dtype=np.float64)
    for mode, rebuild in (("norebuild", False), ("rebuild", True)):
        ms = pw.MolecularSystem.load_file(path)
        ms.make_modular(rebuild=rebuild)
        mols = ms.molecules
        out[f"{mode}_count"] = np.array(len(mols))
        out[f"{mode}_sizes"] = np.array([len(mols[k].elements) for k in sorted(mols)])
        out[f"{mode}_elements"] = np.concatenate([np.asarray(mols[k].elements) for k in sorted(mols)]).astype("U4")
        out[f"{mode}_coords"] = np.concatenate([np.asarray(mols[k].coordinates) for k in sorted(mols)])
    ms = pw.MolecularSystem.load_file(path)
    rb = ms.rebuild_system()
    out["rebuilt_elements"] = np.asarray(rb.system["elements"]).astype("U4")
    out["rebuilt_coords"] = np.asarray(rb.system["coordinates"], dtype=np.float64)
    pw.MolecularSystem.load_file = _load
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"{name}: N={len(out['elements'])} molecules={out['rebuild_count']} / {out['norebuild_count']}", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    pw, util = load_reference()
    jobs = []

    cc3 = load_reference_test_module("test_validate_cc3").system
    jobs.append(("cc3", lambda: analyse(pw, util, cc3["elements"], cc3["coordinates"], "cc3")))
    tw = load_reference_test_module("test_validate_windows")
    for i in range(1, 6):
        c = getattr(tw, f"case_{i}")
        jobs.append((f"win_case{i}", lambda c=c, i=i: analyse(pw, util, c["elements"], c["coordinates"], f"win_case{i}")))
    ta = load_reference_test_module("test_validate_average_diameter")
    for i in range(1, 6):
        c = getattr(ta, f"case_{i}")
        jobs.append((f"avg_case{i}", lambda c=c, i=i: analyse(pw, util, c["elements"], c["coordinates"], f"avg_case{i}", windows=False)))

    inp = os.path.join(REF_ROOT, "examples", "data", "input")

    def md_frames():
        traj = pw.DLPOLY(os.path.join(inp, "HISTORY_singlemol_short"))
        for f in range(traj.no_of_frames):
            ms = traj.get_frames(f, swap_atoms={"he": "H"}, forcefield="opls")[f]
            analyse(pw, util, ms.system["elements"], ms.system["coordinates"], f"md_frame{f:02d}")
    jobs.append(("md", md_frames))

    for nm, fn in (("mibqar", "MIBQAR.pdb"), ("saygor", "SAYGOR.pdb"), ("reymal", "REYMAL.xyz"),
                   ("pudxes", "PUDXES.xyz"), ("yaqhoq", "YAQHOQ.xyz"), ("batvup", "BATVUP.xyz"),
                   ("nuxhiz", "NUXHIZ.xyz")):
        def ex(nm=nm, fn=fn):
            ms = pw.MolecularSystem.load_file(os.path.join(inp, fn))
            analyse(pw, util, ms.system["elements"], ms.system["coordinates"], "ex_" + nm)
        jobs.append(("ex_" + nm, ex))

    for nm, fn, adj in (("mibqar_adj8", "MIBQAR.pdb", 8), ("saygor_adj113", "SAYGOR.pdb", 113)):
        def big(nm=nm, fn=fn, adj=adj):
            ms = pw.MolecularSystem.load_file(os.path.join(inp, fn))
            analyse(pw, util, ms.system["elements"], ms.system["coordinates"], "big_" + nm, adjust=adj)
        jobs.append(("big_" + nm, big))

    for nm, fn in (("pudxes_cell", "PUDXES_no_solvent.pdb"), ("epirur_cell", "EPIRUR_no_solvent.pdb"),
                   ("tatver_cell", "TATVER_no_solvent.pdb")):
        jobs.append(("per_" + nm, lambda nm=nm, fn=fn: periodic(pw, "per_" + nm, os.path.join(inp, fn))))

    for nm, fn, sd in (("pudxes_noise1", "PUDXES_no_solvent.pdb", 20261017), ("pudxes_noise2", "PUDXES_no_solvent.pdb", 20261018),
                       ("epirur_noise", "EPIRUR_no_solvent.pdb", 20261019)):
        jobs.append(("per_" + nm, lambda nm=nm, fn=fn, sd=sd: periodic(pw, "per_" + nm, os.path.join(inp, fn), noise_seed=sd)))

    def writers():
        """Files written by the reference's Output (io_tools.py:349-493) for a fixed system: tests/golden/writer_ref.*"""
        from pywindow._internal.io_tools import Output
        g = np.load(os.path.join(OUT, "md_frame00.npz"))
        system = writer_system(g)
        for ext, kw in (("pdb", dict(atom_ids_key="atom_ids", remarks=["first", "second"], space_group="P21")),
                        ("xyz", dict(atom_ids_key="elements", remarks=["a", "b"]))):
            Output().dump2file(system, os.path.join(OUT, "writer_ref." + ext), override=True, **kw)
    jobs.append(("writers", writers))

    for name, job in jobs:
        if a.only is None or a.only in name:
            job()


if __name__ == "__main__":
    main()
