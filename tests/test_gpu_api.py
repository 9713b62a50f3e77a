"""Drop-in API on the GPU: the reference's own test cases (values quoted from
/root/reference/tests with file:line) through Molecule / MolecularSystem / DLPOLY of pywindow_b200."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _mol(name):
    import pywindow_b200 as pw
    g = load_golden(name)
    ms = pw.MolecularSystem.load_system({"elements": g["elements"], "coordinates": g["coords"]}, name)
    return ms.system_to_molecule(), g


def test_validate_cc3_like_reference():
    """/root/reference/tests/test_validate_cc3.py:353-439 (decimal=3)."""
    mol, _ = _mol("cc3")
    np.testing.assert_almost_equal(mol.calculate_centre_of_mass(), np.array([12.4, 12.4, 12.4]))
    np.testing.assert_almost_equal(mol.calculate_maximum_diameter(), 22.179369990077188, decimal=3)
    np.testing.assert_almost_equal(mol.calculate_average_diameter(), 13.832017514255472, decimal=3)
    np.testing.assert_almost_equal(mol.calculate_pore_diameter(), 5.397020177310022, decimal=3)
    np.testing.assert_almost_equal(mol.calculate_pore_volume(), 82.31154385154417, decimal=3)
    np.testing.assert_almost_equal(mol.calculate_pore_diameter_opt(), 5.397020177310022, decimal=3)
    np.testing.assert_almost_equal(mol.calculate_pore_volume_opt(), 82.31154385154417, decimal=3)
    w = mol.calculate_windows()
    ref = np.array([3.63778746, 3.63562103, 3.63707237, 3.62896512])
    np.testing.assert_almost_equal(np.sort(w), np.sort(ref), decimal=3)
    coms = mol.properties["windows"]["centre_of_mass"]
    ref_c = np.array([[10.77105705, 10.77097707, 14.02893956], [14.01544846, 14.0154126, 14.01539845],
                      [10.77542236, 14.02453217, 10.77546634], [13.92965526, 10.87029764, 10.87034161]])
    np.testing.assert_almost_equal(coms[np.argsort(w)], ref_c[np.argsort(ref)], decimal=2)


def test_full_analysis_matches_individual_calls_and_types():
    mol, g = _mol("md_frame03")
    p = mol.full_analysis()
    assert isinstance(p["no_of_atoms"], int) and isinstance(p["maximum_diameter"]["atom_1"], int)
    assert isinstance(p["average_diameter"], float) and isinstance(p["centre_of_mass"], np.ndarray)
    assert p["windows"]["diameters"].shape == (4,) and p["windows"]["centre_of_mass"].shape == (4, 3)
    mol2, _ = _mol("md_frame03")
    mol2.calculate_centre_of_mass(); mol2.calculate_maximum_diameter(); mol2.calculate_average_diameter()
    mol2.calculate_pore_volume(); mol2.calculate_pore_volume_opt(); mol2.calculate_windows()
    import pywindow_b200 as pw
    assert pw.compare_properties_dict(p, mol2.properties) == (True, "none")
    np.testing.assert_allclose(p["maximum_diameter"]["diameter"], g["maxdim"][2], atol=1e-9)


@pytest.mark.parametrize("case,count", [(1, None), (2, 2), (3, 3), (4, 4), (5, 6)])
def test_validate_windows_like_reference(case, count):
    """/root/reference/tests/test_validate_windows.py:2001-2087."""
    mol, g = _mol(f"win_case{case}")
    w = mol.calculate_windows()
    if count is None:
        assert w is None and mol.properties["windows"] == {"diameters": None, "centre_of_mass": None}
        return
    assert len(w) == count
    np.testing.assert_almost_equal(np.sort(w), np.sort(g["windows_diam"]), decimal=3)


@pytest.mark.parametrize("case,val", list(enumerate([12.38896, 13.36607, 18.10741, 19.23547, 24.03139], start=1)))
def test_validate_average_diameter_like_reference(case, val):
    """/root/reference/tests/test_validate_average_diameter.py:2373-2415."""
    mol, _ = _mol(f"avg_case{case}")
    np.testing.assert_almost_equal(mol.calculate_average_diameter(), val, decimal=3)


def test_inverted_bounds_raise_like_scipy():
    import pywindow_b200 as pw
    names = [n for n in ("ex_yaqhoq", "win_case1", "ex_batvup") if "opt_error" in load_golden(n)]
    for n in names:
        mol, _ = _mol(n)
        with pytest.raises(ValueError):
            mol.calculate_pore_diameter_opt()


def test_trajectory_analysis_matches_per_frame(tmp_path):
    """DLPOLY.analysis (one GPU batch) == Molecule.full_analysis per frame; override / frame lists."""
    import pywindow_b200 as pw
    from pywindow_b200 import synthetic
    gs = [load_golden(f"md_frame{i:02d}") for i in range(5)]
    X = np.stack([g["coords"] for g in gs])
    el = gs[0]["elements"]
    path = tmp_path / "HISTORY"
    synthetic.write_history(path, [e + "1" for e in el], X)
    t = pw.DLPOLY(path)
    t.analysis(frames=[0, 3], forcefield="DLF")
    assert sorted(t.analysis_output) == [0, 3]
    t.analysis(frames="all", forcefield="DLF", ncpus=8)        # skips 0 and 3 (override=False)
    assert sorted(t.analysis_output) == [0, 1, 2, 3, 4]
    for f in range(5):
        ms = t.get_frames(f, forcefield="DLF")[f]
        p = ms.system_to_molecule().full_analysis()
        assert pw.compare_properties_dict(p, t.analysis_output[f]["0"]) == (True, "none"), f
        assert len(p["windows"]["diameters"]) == len(gs[f]["windows_diam"])
    t.save_analysis(tmp_path / "out")
    import json
    js = json.load(open(tmp_path / "out.json"))
    assert len(js) == 5 and len(js["2"]["0"]["windows"]["diameters"]) == 4


def test_host_buffer_entry_point_matches_device_path():
    from pywindow_b200.engine import default_engine
    eng = default_engine()
    gs = [load_golden(f"md_frame{i:02d}") for i in range(4)]
    X = np.stack([g["coords"] for g in gs])
    el = gs[0]["elements"]
    a = eng.analyse_frames(X, el)
    h = eng.analyse_frames_host(X, el, ray_cap=a.ray_cap).numpy().copy()
    np.testing.assert_array_equal(a.host_records(), h)
    # asynchronous entry, two batches in flight on two streams with separate staging state
    import torch
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    states, outs = [{}, {}], []
    Xp = [torch.from_numpy(X.copy()).pin_memory(), torch.from_numpy(X[::-1].copy()).pin_memory()]
    for k in range(2):
        with torch.cuda.stream(streams[k]):
            outs.append(eng.analyse_frames_host(Xp[k], el, ray_cap=a.ray_cap, state=states[k], wait=False))
    for k in range(2):
        streams[k].synchronize()
    np.testing.assert_array_equal(outs[0].numpy(), h)
    np.testing.assert_array_equal(outs[1].numpy(), h[::-1])


def test_xyz_and_pdb_trajectory_analysis_matches_per_frame(tmp_path):
    """XYZ / PDB trajectories through the native frame decoders (pwb_xyz_decode, pwb_pdb_decode) and one GPU
    batch == Molecule.full_analysis of every frame decoded by the Python reader."""
    import pywindow_b200 as pw
    from pywindow_b200 import synthetic
    gs = [load_golden(f"md_frame{i:02d}") for i in range(4)]
    X = np.stack([g["coords"] for g in gs])
    el = gs[0]["elements"]
    synthetic.write_xyz_trajectory(tmp_path / "t.xyz", el, X)
    synthetic.write_pdb_trajectory(tmp_path / "t.pdb", el, X)
    for t in (pw.XYZ(tmp_path / "t.xyz"), pw.PDB(tmp_path / "t.pdb")):
        assert t.no_of_frames == 4
        t.analysis(frames="all")
        assert sorted(t.analysis_output) == [0, 1, 2, 3]
        for f in range(4):
            p = t.get_frames(f)[f].system_to_molecule().full_analysis()
            assert pw.compare_properties_dict(p, t.analysis_output[f]["0"]) == (True, "none"), (type(t).__name__, f)
            assert len(p["windows"]["diameters"]) == 4


def test_pipelined_trajectory_batches_equal_one_batch(tmp_path):
    """Trajectory.analysis over several pipeline batches (short head batch, two pinned slots used alternately,
    records leaving the slots on the helper thread, a last partial batch) gives exactly the records of ONE device
    batch over the same frames, in frame order."""
    import pywindow_b200 as pw
    from pywindow_b200 import synthetic
    from pywindow_b200.engine import default_engine
    gs = [load_golden(f"md_frame{i:02d}") for i in range(5)]
    seeds = np.stack([g["coords"] for g in gs])
    el = [str(e) for e in gs[0]["elements"]]
    F = 5000
    X = synthetic.noisy_frames(seeds, 0, F)
    path = tmp_path / "HISTORY"
    synthetic.write_history_native(str(path), [e + "1" for e in el], X, compact=True)
    t = pw.DLPOLY(path)
    t.batch_frames = 700                                     # head 87 + 7 full batches + 13 frames
    t.analysis(forcefield="DLF")
    tab, keys = t.analysis_output.record_table()
    assert tab.shape[0] == F and list(keys.frames) == list(range(F))
    Xd = t._decode_coordinates(list(range(F)))               # what the file holds (4 significant decimals)
    ref = default_engine().analyse_frames(Xd, el).host_records()
    np.testing.assert_array_equal(tab, ref)


def test_bounce_buffer_readback_is_exact():
    """distributed._device_to_host (pinned bounce buffers + helper threads) == Tensor.cpu() for float64 and int64
    tensors larger than the buffers, including a ragged last chunk."""
    import torch
    from pywindow_b200.distributed import _device_to_host, _BOUNCE_DOUBLES
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn((3 * _BOUNCE_DOUBLES + 12345) // 92, 92, dtype=torch.float64, device="cuda", generator=g)
    np.testing.assert_array_equal(_device_to_host(a), a.cpu().numpy())
    k = torch.randint(-5, 1 << 40, (2, _BOUNCE_DOUBLES + 77), dtype=torch.int64, device="cuda", generator=g)
    out = _device_to_host(k)
    assert out.dtype == np.int64 and out.shape == (2, _BOUNCE_DOUBLES + 77)
    np.testing.assert_array_equal(out, k.cpu().numpy())
    small = torch.arange(10, dtype=torch.float64, device="cuda")
    np.testing.assert_array_equal(_device_to_host(small), np.arange(10.0))
