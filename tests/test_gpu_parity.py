"""GPU parity: the CUDA path (through the C ABI, via pywindow_b200.engine) against the golden
fixtures recorded from the reference (tests/golden, oracle/make_golden.py) and the CPU oracle.

Protocol: (1) stage-wise / teacher-forced with the reference's pore centre --
deterministic stages to 1e-9, ray masks / DBSCAN labels / window counts exact, window diameters
1e-3 A; (2) free-running -- window count exact, optimised pore diameter within the reference's own
noise floor."""
import numpy as np
import pytest

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu

NAMES = golden_names(("cc3", "win_", "md_", "ex_", "avg_"))


@pytest.fixture(scope="module")
def eng():
    from pywindow_b200.engine import Engine
    return Engine(0)


@pytest.fixture(scope="module")
def goldens():
    return {n: load_golden(n) for n in NAMES}


@pytest.fixture(scope="module")
def forced(eng, goldens):
    """One heterogeneous batch of every golden molecule, pore centre forced to the reference's."""
    mols, centres = [], []
    for n in NAMES:
        g = goldens[n]
        mols.append((g["elements"], g["coords"]))
        centres.append(g["fw_centre"] if "fw_centre" in g else g["com"])
    res = eng.analyse_molecules(mols, centre_override=np.array(centres))
    return res, res.host_records()


@pytest.fixture(scope="module")
def free(eng, goldens):
    mols = [(goldens[n]["elements"], goldens[n]["coords"]) for n in NAMES]
    res = eng.analyse_molecules(mols)
    return res, res.host_records()


def test_scalars(forced, goldens):
    from pywindow_b200 import _lib as L
    _, rec = forced
    for i, n in enumerate(NAMES):
        g, r = goldens[n], rec[i]
        assert int(r[L.REC_NATOMS]) == len(g["elements"])
        np.testing.assert_allclose(r[L.REC_COM:L.REC_COM + 3], g["com"], atol=1e-9, err_msg=n)
        assert abs(r[L.REC_MAXD] - g["maxdim"][2]) < 1e-9, n
        assert abs(r[L.REC_PORE_D] - g["pore"][0]) < 1e-9, n
        assert abs(r[L.REC_AVGD] - float(g["avg_diameter"])) < 1e-9, n
        X = g["coords"]
        if n.startswith("md_"):     # tie-free inputs: atom indices exact
            assert (int(r[L.REC_MAXD_A1]), int(r[L.REC_MAXD_A2])) == (int(g["maxdim"][0]), int(g["maxdim"][1])), n
            assert int(r[L.REC_PORE_ATOM]) == int(g["pore"][1]), n
        else:                       # crystal symmetry ties: index must attain the optimum
            import pw_oracle as O
            rv = O.vdw_of(g["elements"])
            i1, i2 = int(r[L.REC_MAXD_A1]), int(r[L.REC_MAXD_A2])
            assert i1 <= i2
            assert abs(np.linalg.norm(X[i1] - X[i2]) + rv[i1] + rv[i2] - g["maxdim"][2]) < 1e-9, n
            assert abs(2 * O.clearances(rv, X, g["com"])[int(r[L.REC_PORE_ATOM])] - g["pore"][0]) < 1e-9, n


def test_forced_opt_fields(forced, goldens):
    """With a forced centre the 'optimised' pore diameter is pore_diameter(centre)."""
    from pywindow_b200 import _lib as L
    import pw_oracle as O
    _, rec = forced
    for i, n in enumerate(NAMES):
        g = goldens[n]
        if "fw_centre" not in g:
            continue
        d, a = O.pore_diameter(O.vdw_of(g["elements"]), g["coords"], g["fw_centre"])
        assert abs(rec[i][L.REC_OPT_D] - d) < 1e-9, n
        assert abs(rec[i][L.REC_OPT_D] - g["opt"][0]) < 1e-9, n


def test_ray_stage(forced, goldens):
    from pywindow_b200 import _lib as L
    res, rec = forced
    ck, cp, ca = res.clear_k.cpu().numpy(), res.clear_pos.cpu().numpy(), res.clear_a2.cpu().numpy()
    for i, n in enumerate(NAMES):
        g = goldens[n]
        if "ray_clear" not in g:
            continue
        assert int(rec[i][L.REC_NRAYS]) == g["ray_clear"].shape[0], n
        want = np.nonzero(g["ray_clear"])[0]
        S = int(rec[i][L.REC_NCLEAR])
        assert S == len(want), n
        assert (ck[i, :S] == want).all(), n
        recs = g["ray_records"]
        if S:
            np.testing.assert_allclose(ca[i, :S], recs[:, 1], atol=1e-9, err_msg=n)
            # column 0 = |chunk*pos|: recover the step index from the record
            R = np.linalg.norm(recs[:, 5:8], axis=1)
            chunks = np.floor_divide(R, 1.0)
            pos = np.rint(recs[:, 0] / (R / chunks)).astype(int)
            assert (cp[i, :S] == pos).all(), n
        assert bool(int(rec[i][L.REC_STATUS]) & L.ST_NO_CLEAR_RAY) == (S == 0), n


def test_clusters(forced, goldens):
    from pywindow_b200 import _lib as L
    res, rec = forced
    lab = res.labels.cpu().numpy()
    for i, n in enumerate(NAMES):
        g = goldens[n]
        if "labels" not in g:
            continue
        S = int(rec[i][L.REC_NCLEAR])
        assert abs(rec[i][L.REC_EPS] - float(g["eps"])) < 1e-9, n
        assert (lab[i, :S] == g["labels"]).all(), n
        assert int(rec[i][L.REC_NWIN]) == g["win_diam"].shape[0], n


def test_windows_teacher_forced(forced, goldens):
    from pywindow_b200.engine import record_windows
    _, rec = forced
    worst = 0.0
    for i, n in enumerate(NAMES):
        g = goldens[n]
        if "has_windows" not in g:
            continue
        w = record_windows(rec[i], warn=False)
        if not bool(g["has_windows"]):
            assert w is None, n
            continue
        assert w is not None and w[0].shape == g["windows_diam"].shape, n     # count incl. order
        worst = max(worst, np.abs(w[0] - g["windows_diam"]).max())
        np.testing.assert_allclose(w[0], g["windows_diam"], atol=1e-3, err_msg=n)
        np.testing.assert_allclose(w[1], g["windows_com"], atol=5e-3, err_msg=n)
    print("worst teacher-forced window diameter difference: %.3e A" % worst)


def test_free_running(free, goldens):
    """No forcing: window COUNT must equal the reference's on every molecule.  The optimised pore diameter is held
    to 1e-3 A where the reference's own tests pin it (the CC3 crystal structure, whose optimiser starts at the
    optimum) and, on the 20 MD frames, to the reference's MEASURED self-agreement under
    1e-12 A jitter (tests/golden/noise_floor.npz; the 220-frame version of this check, incl. free-running window
    diameters, is tests/test_noise_floor.py)."""
    from pywindow_b200 import _lib as L
    from pywindow_b200.engine import record_windows
    _, rec = free
    d = {}
    for i, n in enumerate(NAMES):
        g = goldens[n]
        if "opt_error" in g:
            assert int(rec[i][L.REC_STATUS]) & L.ST_OPT_BOUNDS, n
            continue
        d[n] = abs(rec[i][L.REC_OPT_D] - g["opt"][0])
        if "has_windows" in g:
            w = record_windows(rec[i], warn=False)
            if not bool(g["has_windows"]):
                assert w is None, n
            else:
                assert w is not None and len(w[0]) == len(g["windows_diam"]), n
    pinned = {n: v for n, v in d.items() if n.startswith("cc3")}      # tests/test_validate_cc3.py:398, test_molecular.py:4611
    md = np.array([v for n, v in d.items() if n.startswith("md_")])
    other = {n: v for n, v in d.items() if n.startswith(("ex_", "win_", "avg_"))}
    print("free-running pore_opt: pinned worst %.2e; md %d/%d within 1e-3, worst %.2e; examples worst %.2e (%s)" % (
        max(pinned.values()), (md <= 1e-3).sum(), len(md), md.max(), max(other.values()), max(other, key=other.get)))
    assert max(pinned.values()) <= 1e-3, pinned
    floor = load_golden("noise_floor")
    ref = np.abs(floor["jit_opt_d"] - floor["base_opt_d"])[:20]          # the same 20 MD frames, reference vs itself
    assert (md > 1e-3).mean() <= (ref > 1e-3).mean() + 0.10
    assert md.max() <= 1.5 * np.abs(floor["jit_opt_d"] - floor["base_opt_d"]).max()
    assert max(other.values()) <= 1.5 * np.abs(floor["jit_opt_d"] - floor["base_opt_d"]).max()


def test_batch_vs_single_and_shared_types(eng, goldens):
    """Same frames as a shared-type trajectory batch and as single items give identical records."""
    names = [n for n in NAMES if n.startswith("md_")][:6]
    X = np.stack([goldens[n]["coords"] for n in names])
    el = goldens[names[0]]["elements"]
    a = eng.analyse_frames(X, el).host_records()
    for i, n in enumerate(names):
        b = eng.analyse_frames(X[i:i + 1], el).host_records()[0]
        np.testing.assert_array_equal(a[i], b)
    c = eng.analyse_molecules([(el, x) for x in X]).host_records()
    np.testing.assert_array_equal(a, c)


BIG = golden_names(("big_",))


@pytest.mark.parametrize("name", BIG)
def test_dense_ray_sampling(eng, name):
    """BASELINE config 5: find_windows / find_average_diameter with adjust >> 1 (7e3 and 1e5 rays),
    teacher-forced with the reference's pore centre: ray count, clear-ray mask, eps, DBSCAN labels,
    window count exact; diameters 1e-3 A."""
    from pywindow_b200 import _lib as L
    from pywindow_b200.engine import record_windows
    g = load_golden(name)
    adj = float(g["adjust"])
    res = eng.analyse_molecules([(g["elements"], g["coords"])], centre_override=g["fw_centre"][None], adjust=adj)
    rec = res.host_records()[0]
    P = g["ray_clear"].shape[0]
    assert int(rec[L.REC_NRAYS]) == P
    S = int(rec[L.REC_NCLEAR])
    want = np.nonzero(g["ray_clear"])[0]
    assert S == len(want)
    assert (res.clear_k.cpu().numpy()[0, :S] == want).all()
    np.testing.assert_allclose(res.clear_a2.cpu().numpy()[0, :S], g["ray_a2"], atol=1e-9)
    assert abs(rec[L.REC_EPS] - float(g["eps"])) < 1e-9
    assert (res.labels.cpu().numpy()[0, :S] == g["labels"]).all()
    assert abs(rec[L.REC_AVGD] - float(g["avg_diameter"])) < 1e-9
    w = record_windows(rec, warn=False)
    assert w is not None and w[0].shape == g["windows_diam"].shape
    np.testing.assert_allclose(w[0], g["windows_diam"], atol=1e-3)
    np.testing.assert_allclose(w[1], g["windows_com"], atol=5e-3)


def test_edge_cases(eng):
    """Empty batch, a large synthetic shell (1 800 atoms: reduced warps per CTA), tiny molecules."""
    from pywindow_b200 import _lib as L
    import pw_oracle as O
    g = load_golden("md_frame00")
    assert eng.analyse_frames(np.zeros((0, 168, 3)), g["elements"]).host_records().shape[0] == 0
    rng = np.random.default_rng(5)
    n = 1800
    u = rng.normal(size=(n, 3))
    u /= np.linalg.norm(u, axis=1)[:, None]
    X = u * rng.uniform(14.0, 16.0, size=(n, 1))
    keep = (np.abs(u[:, 2]) < 0.93)                       # two polar openings
    X = X[keep]
    el = np.array(["C", "H", "N", "O"])[rng.integers(0, 4, size=X.shape[0])]
    rec = eng.analyse_molecules([(el, X)]).host_records()[0]
    r, m = O.vdw_of(el), O.masses_of(el)
    assert abs(rec[L.REC_MAXD] - O.max_dim(r, X)[2]) < 1e-9
    assert abs(rec[L.REC_AVGD] - O.find_average_diameter(r, m, X)) < 1e-9
    w = O.find_windows(r, m, X, centre=rec[L.REC_OPT_C:L.REC_OPT_C + 3])
    nw = 0 if w is None else len(w[0])
    assert int(rec[L.REC_NWIN]) == nw and nw >= 2
    got = np.sort(rec[L.REC_WIN0:L.REC_WIN0 + 4 * nw].reshape(nw, 4)[:, 0])
    np.testing.assert_allclose(got, np.sort(w[0]), atol=1e-3)
    # water-sized molecule: no cavity, COM inside an atom -> inverted optimiser bounds flag, no crash
    tiny = eng.analyse_molecules([(np.array(["O", "H", "H"]), np.array([[0, 0, 0.0], [0.96, 0, 0], [-0.24, 0.93, 0]]))])
    assert int(tiny.host_records()[0][L.REC_STATUS]) & L.ST_OPT_BOUNDS


def test_seeded_synthetic_frames_against_oracle(eng):
    """The bench workload itself (seeded noisy CC3 MD frames) on the GPU vs the CPU oracle: every
    deterministic field to 1e-9, window count exact, window diameters 1e-3 A with the oracle started
    from the GPU's pore centre (the only chaotic quantity, compared separately)."""
    import pw_oracle as O
    from pywindow_b200 import _lib as L, synthetic
    from pywindow_b200.engine import record_windows
    seeds = np.stack([load_golden(f"md_frame{i:02d}")["coords"] for i in range(20)])
    el = load_golden("md_frame00")["elements"]
    X = synthetic.noisy_frames(seeds, 1000, 24)
    rec = eng.analyse_frames(X, el).host_records()
    r, m = O.vdw_of(el), O.masses_of(el)
    near = 0
    for f in range(X.shape[0]):
        com = O.center_of_mass(m, X[f])
        np.testing.assert_allclose(rec[f, L.REC_COM:L.REC_COM + 3], com, atol=1e-9)
        i1, i2, D = O.max_dim(r, X[f])
        assert abs(rec[f, L.REC_MAXD] - D) < 1e-9 and (int(rec[f, L.REC_MAXD_A1]), int(rec[f, L.REC_MAXD_A2])) == (i1, i2)
        pd_, pa = O.pore_diameter(r, X[f], com)
        assert abs(rec[f, L.REC_PORE_D] - pd_) < 1e-9 and int(rec[f, L.REC_PORE_ATOM]) == pa
        assert abs(rec[f, L.REC_AVGD] - O.find_average_diameter(r, m, X[f])) < 1e-9
        c = rec[f, L.REC_OPT_C:L.REC_OPT_C + 3]
        od, oa = O.pore_diameter(r, X[f], c)
        assert abs(rec[f, L.REC_OPT_D] - od) < 1e-9 and int(rec[f, L.REC_OPT_ATOM]) == oa
        near += abs(O.opt_pore_diameter(r, X[f], com)[0] - od) <= 1e-3
        w = O.find_windows(r, m, X[f], centre=c)
        g = record_windows(rec[f], warn=False)
        assert w is not None and g is not None and len(w[0]) == len(g[0]) == 4
        np.testing.assert_allclose(g[0], w[0], atol=1e-3)
        np.testing.assert_allclose(g[1], w[1], atol=5e-3)
    # free-running pore centre, GPU vs ORACLE (two independent draws from the same chaotic optimiser; the
    # reference-anchored version of this check is tests/test_noise_floor.py)
    assert near >= 0.6 * X.shape[0]


def test_path_scan_includes_the_ray_end_point(eng):
    """vector_analysis evaluates steps 0..chunks INCLUSIVE (utilities.py:1112).  On sparse lopsided shells with an
    off-centre ray origin the narrowest point of many clear rays IS the end point, so dropping it changes
    2*min clearance, the arg-min step and sometimes the clear flag.  Clear-ray list, arg-min steps and
    clearances must equal the oracle's stepped scan exactly."""
    import pw_oracle as O
    hit_last = 0
    for seed in (1, 7, 11):
        rng = np.random.default_rng(seed)
        n = int(rng.integers(20, 60))
        u = rng.normal(size=(n, 3))
        u /= np.linalg.norm(u, axis=1)[:, None]
        X = u * (rng.uniform(5, 9, size=(n, 1)) * np.array([1 + 0.6 * (u[:, 0] > 0)]).T)
        el = np.array(["C", "H", "N"])[rng.integers(0, 3, size=n)]
        r, m = O.vdw_of(el), O.masses_of(el)
        centre = O.center_of_mass(m, X) + rng.uniform(-2, 2, 3)
        det = {}
        O.find_windows(r, m, X, centre=centre, detail=det)
        res = eng.analyse_molecules([(el, X)], centre_override=centre[None])
        rec = res.host_records()[0]
        S = int(rec[17])                                    # PWB_REC_NCLEAR
        want = np.nonzero(det["clear"])[0]
        assert S == len(want) and (res.clear_k.cpu().numpy()[0, :S] == want).all(), seed
        recs = det["records"]
        np.testing.assert_allclose(res.clear_a2.cpu().numpy()[0, :S], recs[:, 1], atol=1e-9)
        Rn = np.linalg.norm(recs[:, 5:8], axis=1)
        chunks = np.floor_divide(Rn, 1.0)
        pos = np.rint(recs[:, 0] / (Rn / chunks)).astype(int)
        assert (res.clear_pos.cpu().numpy()[0, :S] == pos).all(), seed
        hit_last += int((pos == chunks).sum())
    assert hit_last > 20, "the cases must exercise the last path step"


def test_find_windows_without_pore_optimisation_ignores_inverted_bounds(eng):
    """find_windows(pore_opt=False) never calls opt_pore_diameter in the reference (utilities.py:1386-1393): a
    molecule whose COM sits inside an atom's vdW sphere still gets its ray analysis instead of scipy's
    inverted-bounds ValueError."""
    import pw_oracle as O
    from pywindow_b200 import utilities as U
    g = load_golden("md_frame00")
    el = np.concatenate([g["elements"], np.array(["C"])])
    r0, m0 = O.vdw_of(g["elements"]), O.masses_of(g["elements"])
    X = np.concatenate([g["coords"], O.center_of_mass(m0, g["coords"])[None] + 0.05])   # an atom on the COM
    with pytest.raises(ValueError):
        U.find_windows(el, X)
    got = U.find_windows(el, X, pore_opt=False)
    want = O.find_windows(O.vdw_of(el), O.masses_of(el), X, pore_opt=False)
    assert (got is None) == (want is None)
    if want is not None:
        assert len(got[0]) == len(want[0])
        np.testing.assert_allclose(got[0], want[0], atol=1e-3)


def test_opt_pore_diameter_bounds_and_start_point(eng):
    """opt_pore_diameter(bounds=..., com=...) (utilities.py:400-426): custom start point, custom / one-sided /
    absent bounds, degenerate and inverted boxes -- against the oracle (same scipy call as the reference)."""
    import pw_oracle as O
    from pywindow_b200 import utilities as U
    g = load_golden("md_frame05")
    el, X = g["elements"], g["coords"]
    r, m = O.vdw_of(el), O.masses_of(el)
    com = O.center_of_mass(m, X)
    start = com + np.array([0.4, -0.3, 0.2])
    cases = [
        dict(com=start),                                                                   # box around the start point
        dict(bounds=tuple((c - 1.0, c + 1.0) for c in com)),                               # custom box, default start
        dict(bounds=tuple((c - 0.7, c + 1.2) for c in start), com=start),
        dict(bounds=((com[0] - 0.5, None), (None, com[1] + 0.5), (None, None)), com=com),  # one-sided / unbounded
    ]
    close = 0
    for kw in cases:
        d, a, c = U.opt_pore_diameter(el, X, **kw)
        x0 = kw.get("com", com)
        od, oa, oc = O.opt_pore_diameter(r, X, np.asarray(x0), kw.get("bounds"))
        # the returned diameter / atom are pore_diameter at the returned centre, and the optimiser did not lose
        pd_, pa = O.pore_diameter(r, X, c)
        assert abs(d - pd_) < 1e-9 and a == pa
        assert d >= O.pore_diameter(r, X, np.asarray(x0))[0] - 1e-12
        if "bounds" in kw:
            for k, (lo, hi) in enumerate(kw["bounds"]):
                assert (lo is None or c[k] >= lo) and (hi is None or c[k] <= hi)
        else:
            pr = O.pore_diameter(r, X, np.asarray(x0))[0] / 2
            assert (np.abs(c - x0) <= pr + 1e-12).all()
        assert abs(d - od) < 5e-2, (kw, d, od)       # chaotic stopping point: the reference's own noise floor
        close += abs(d - od) <= 1e-3
    assert close >= 2
    # degenerate box: nothing to optimise, the answer is pore_diameter(start) exactly
    d, a, c = U.opt_pore_diameter(el, X, bounds=tuple((v, v) for v in start), com=start)
    assert np.array_equal(c, start) and abs(d - O.pore_diameter(r, X, start)[0]) < 1e-12
    with pytest.raises(ValueError):
        U.opt_pore_diameter(el, X, bounds=((1.0, -1.0), (0.0, 1.0), (0.0, 1.0)))


def test_molecule_larger_than_shared_memory(eng):
    """A 5 600-atom shell (the per-atom tiles of k_prepare / k_rays / k_windows no longer fit 227 KB of shared memory
    and live in the global scratch region): every deterministic quantity against the oracle to 1e-9, windows found
    from the GPU's pore centre to 1e-3 A -- the reference has no size limit, neither has the GPU path."""
    from pywindow_b200 import _lib as L
    from pywindow_b200.engine import record_windows
    import pw_oracle as O
    rng = np.random.default_rng(17)
    n = 6200
    u = rng.normal(size=(n, 3))
    u /= np.linalg.norm(u, axis=1)[:, None]
    X = u * rng.uniform(46.0, 48.5, size=(n, 1))
    X = X[np.abs(u[:, 2]) < 0.955]                         # two polar openings
    el = np.array(["C", "H", "N", "O"])[rng.integers(0, 4, size=X.shape[0])]
    assert X.shape[0] > 5000
    res = eng.analyse_molecules([(el, X)])
    rec = res.host_records()[0]
    assert int(rec[L.REC_STATUS]) & ~(L.ST_WIN_NEGATIVE | L.ST_WIN_DROPPED) == 0
    r, m = O.vdw_of(el), O.masses_of(el)
    com = O.center_of_mass(m, X)
    np.testing.assert_allclose(rec[L.REC_COM:L.REC_COM + 3], com, atol=1e-9)
    i1, i2, D = O.max_dim(r, X)
    assert abs(rec[L.REC_MAXD] - D) < 1e-9
    pd_, pa = O.pore_diameter(r, X, com)
    assert abs(rec[L.REC_PORE_D] - pd_) < 1e-9 and int(rec[L.REC_PORE_ATOM]) == pa
    assert abs(rec[L.REC_AVGD] - O.find_average_diameter(r, m, X)) < 1e-9
    c = rec[L.REC_OPT_C:L.REC_OPT_C + 3]
    assert abs(rec[L.REC_OPT_D] - O.pore_diameter(r, X, c)[0]) < 1e-9
    det = {}
    w = O.find_windows(r, m, X, centre=c, detail=det)
    S = int(rec[L.REC_NCLEAR])
    assert int(rec[L.REC_NRAYS]) == det["P"] and S == int(det["clear"].sum())
    assert (res.clear_k.cpu().numpy()[0, :S] == np.nonzero(det["clear"])[0]).all()
    assert (res.labels.cpu().numpy()[0, :S] == det["labels"]).all()
    g = record_windows(rec, warn=False)
    assert w is not None and g is not None and len(g[0]) == len(w[0]) >= 2
    np.testing.assert_allclose(g[0], w[0], atol=1e-3)
    np.testing.assert_allclose(g[1], w[1], atol=5e-3)
    # the same molecule in a batch with small ones (per-item tiles sized by the largest item)
    small = load_golden("md_frame00")
    both = eng.analyse_molecules([(small["elements"], small["coords"]), (el, X)]).host_records()
    one = eng.analyse_molecules([(small["elements"], small["coords"])]).host_records()[0]
    np.testing.assert_array_equal(both[1], rec)
    np.testing.assert_array_equal(both[0], one)


def test_survivor_lists_of_the_pore_search_are_exact(eng):
    """k_opt_pore evaluates its objective over the atoms that CAN attain the minimum within 0.5 A of a reference point
    (1-Lipschitz bound, csrc/pwb_device.cuh: group_build_list).  With the lists switched off (every atom at every
    point) the records must be bit-identical: MD frames, the same frames 150 A from the origin (forward-difference
    steps 100 x larger), user bounds / start points (the second instantiation of the kernel), a 12-atom molecule and a
    5 900-atom shell whose list overflows."""
    from pywindow_b200 import _lib as L, synthetic
    from pywindow_b200.engine import Engine
    import torch
    gs = [load_golden(f"md_frame{i:02d}") for i in range(4)]
    seeds = np.stack([g["coords"] for g in gs])
    el = [str(e) for e in gs[0]["elements"]]
    X = synthetic.noisy_frames(seeds, 0, 3000)
    rng = np.random.default_rng(23)
    n = 6200
    u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
    shell = (u * rng.uniform(46.0, 48.5, size=(n, 1)))[np.abs(u[:, 2]) < 0.955]
    shell_el = list(np.array(["C", "H", "N", "O"])[rng.integers(0, 4, size=shell.shape[0])])
    tiny = (["C"] * 6 + ["H"] * 6, np.concatenate([2.0 * np.eye(3), -2.0 * np.eye(3), 3.1 * np.eye(3), -3.1 * np.eye(3)]))
    com = X[:200].mean(axis=1)
    start = com + rng.normal(0, 0.3, size=com.shape)
    bounds = np.stack([np.stack([start[:, i] - rng.uniform(0.2, 2.0, 200), start[:, i] + rng.uniform(0.2, 2.0, 200)], axis=1)
                       for i in range(3)], axis=1)                      # [200, 3, 2]
    e2 = Engine(0)
    try:
        out = {}
        for no_list in (0, 1):
            e2.set_option(L.OPTION_OPT_NO_LIST, no_list)
            stages = L.STAGE_SCALARS | L.STAGE_OPT
            out[no_list] = [
                e2.analyse_frames(X, el, stages=stages).host_records(),
                e2.analyse_frames(X[:500] + np.array([150.0, -90.0, 40.0]), el, stages=stages).host_records(),
                e2.analyse_frames(X[:200], el, stages=stages, opt_start=start, opt_bounds=bounds.reshape(200, 6)).host_records(),
                e2.analyse_molecules([tiny, (shell_el, shell)], stages=stages).host_records(),
            ]
        for a, b in zip(out[0], out[1]):
            assert (a[:, L.REC_OPT_NFEV] > 0).all()
            np.testing.assert_array_equal(a, b)
        assert out[0][0][:, L.REC_OPT_NFEV].mean() > 100          # the searches really ran
    finally:
        e2.close()
        torch.cuda.synchronize()


def test_more_than_65535_atoms_per_molecule_is_refused(eng):
    """Atom indices inside the kernels are 16 bit (near lists, survivor lists): a larger molecule must fail loudly."""
    from pywindow_b200 import _lib as L
    rng = np.random.default_rng(5)
    n = 70000
    u = rng.normal(size=(n, 3))
    X = u / np.linalg.norm(u, axis=1)[:, None] * 160.0
    with pytest.raises(L.PwbLibraryError, match="65535"):
        eng.analyse_molecules([(["C"] * n, X)])
