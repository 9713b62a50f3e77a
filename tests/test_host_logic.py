"""CPU tests of the host side: C-ABI symbols, trajectory file mapping / decoding, frame selection,
sharding arithmetic, atom-key deciphering, comparator.  No GPU compute."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden


def test_abi_exports_every_declared_symbol():
    """libpwb.so loads and exports every function include/pwb.h declares."""
    hdr = open(os.path.join(ROOT, "include", "pwb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pwb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    so = os.path.join(ROOT, "pywindow_b200", "csrc", "libpwb.so")
    assert os.path.exists(so), "run python __graft_entry__.py"
    lib = ctypes.CDLL(so)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    lib.pwb_abi_version.restype = ctypes.c_int
    assert lib.pwb_abi_version() == 4
    from pywindow_b200 import _lib
    assert set(_lib.EXPORTS) <= declared


def test_workspace_layout_is_consistent():
    from pywindow_b200 import _lib
    L = _lib.load()
    lay = _lib.WsLayout()
    L.pwb_workspace_layout(10, 1680, 168, 896, 16, ctypes.byref(lay))
    assert lay.total == L.pwb_workspace_bytes(10, 1680, 168, 896, 16)
    assert lay.atom_scratch == lay.opt_hist                    # no atom scratch for molecules that fit shared memory
    assert lay.total - lay.opt_hist >= 10 * 60 * 8             # the (s, y) rings of the pore-centre searches
    big = _lib.WsLayout()
    L.pwb_workspace_layout(2, 12000, 6000, 1024, 16, ctypes.byref(big))          # atoms beyond shared memory
    assert big.total > big.opt_hist > big.atom_scratch >= big.win_k
    offs = [lay.atoms, lay.atom_idx, lay.item_d, lay.item_i, lay.clear_k, lay.clear_pos, lay.clear_a2, lay.labels, lay.win_k]
    assert offs == sorted(offs) and all(o % 256 == 0 for o in offs)


def test_engine_refuses_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pywindow_b200.engine import Engine, PwbNoDeviceError
    with pytest.raises(PwbNoDeviceError):
        Engine()


@pytest.fixture()
def md(tmp_path):
    from pywindow_b200 import synthetic
    g = [load_golden(f"md_frame{i:02d}") for i in range(3)]
    X = np.stack([x["coords"] for x in g])
    el = g[0]["elements"]
    return tmp_path, el, X, synthetic


def test_dlpoly_mapping_and_decode(md):
    tmp, el, X, syn = md
    from pywindow_b200.trajectory import DLPOLY
    p = tmp / "HISTORY"
    lat = np.diag([30.0, 31.0, 32.0])
    syn.write_history(p, list(el), X, lattice=lat)
    t = DLPOLY(p)
    assert t.no_of_frames == 3 and t.no_of_atoms == len(el) and t.periodic_boundary == "parallelepiped"
    fr = t._decode_frame(t._frame_bytes(1))
    np.testing.assert_allclose(fr["coordinates"], X[1], rtol=1e-10)
    np.testing.assert_allclose(fr["lattice"], lat)
    assert list(fr["atom_ids"]) == list(el)
    np.testing.assert_allclose(t._decode_coordinates([0, 2]), X[[0, 2]], rtol=1e-10)
    ms = t.get_frames(2)[2]
    assert ms.system["coordinates"].shape == (len(el), 3)


def test_xyz_and_pdb_trajectories(md):
    tmp, el, X, syn = md
    from pywindow_b200.trajectory import PDB, XYZ
    syn.write_xyz_trajectory(tmp / "t.xyz", el, X)
    t = XYZ(tmp / "t.xyz")
    assert t.no_of_frames == 3
    np.testing.assert_allclose(t._decode_coordinates([1])[0], X[1], atol=1e-7)
    syn.write_pdb_trajectory(tmp / "t.pdb", el, X, unit_cell=[30, 30, 30, 90, 90, 90])
    t = PDB(tmp / "t.pdb")
    assert t.no_of_frames == 3
    fr = t._decode_frame(t._frame_bytes(2))
    np.testing.assert_allclose(fr["coordinates"], X[2], atol=6e-4)
    assert list(fr["elements"]) == list(el) and fr["unit_cell"].shape == (6,)


def test_frame_selection_rules(md):
    tmp, el, X, syn = md
    from pywindow_b200.trajectory import DLPOLY, _FunctionError
    syn.write_history(tmp / "H", list(el), X)
    t = DLPOLY(tmp / "H")
    assert t._select("all") == [0, 1, 2] and t._select(1) == [1] and t._select([2, 0]) == [2, 0]
    assert t._select((0, 2)) == [0, 1]          # the reference raises here; superset
    for bad in ("some", [0, "a"], (0, 1, 2)):
        with pytest.raises(_FunctionError):
            t._select(bad)


def test_shard_ranges_cover_everything():
    from pywindow_b200.distributed import shard_range
    for n in (0, 1, 7, 8, 100, 100001):
        for size in (1, 2, 3, 8):
            r = [shard_range(n, k, size) for k in range(size)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(size - 1))
            assert max(b - a for a, b in r) <= -(-n // size)


def test_atom_key_deciphering_and_comparator():
    from pywindow_b200 import utilities as U
    assert U.decipher_atom_key("C1", "DLF") == "C" and U.decipher_atom_key("Cl12", "dl_f") == "Cl"
    assert U.decipher_atom_key("ca", "opls") == "C" and U.decipher_atom_key("ni", "OPLSAA") == "N"
    with pytest.raises(U._AtomKeyConflictError):
        U.decipher_atom_key("he", "opls")
    with pytest.raises(U._ForceFieldError):
        U.decipher_atom_key("C1", "amber")
    a = {"no_of_atoms": 3, "pore_diameter": {"diameter": 1.0, "atom": 2}, "windows": {"diameters": None, "centre_of_mass": None}}
    b = {"no_of_atoms": 3, "pore_diameter": {"diameter": 1.0 + 1e-12, "atom": 2}, "windows": {"diameters": None, "centre_of_mass": None}}
    assert U.compare_properties_dict(a, b) == (True, "none")
    b["pore_diameter"]["atom"] = 1
    assert U.compare_properties_dict(a, b) == (False, "pore_diameter.atom")
    lat = U.unit_cell_to_lattice_array([10, 11, 12, 80, 95, 100])
    np.testing.assert_allclose(U.lattice_array_to_unit_cell(lat), [10, 11, 12, 80, 95, 100])


def test_native_history_decoder_is_bit_exact(tmp_path):
    """pwb_history_decode (C++, threaded) against the Python per-frame decoder on awkward number
    formats, velocities/forces records (keytrj 2) and a lattice block (imcon 3)."""
    from pywindow_b200.trajectory import DLPOLY
    rng = np.random.default_rng(3)
    n, F = 37, 11
    fmts = ["%.17E", "%12.4E", "%20.10f", "%.3f", "%.17g", "%25.16e"]
    path = tmp_path / "HISTORY"
    X = rng.normal(0, 30, size=(F, n, 3)) * 10.0 ** rng.integers(-12, 6, size=(F, n, 1))
    lat = rng.normal(0, 20, size=(F, 3, 3))
    with open(path, "w") as fh:
        fh.write("title\n         2         3        %d\n" % n)
        for f in range(F):
            fh.write("timestep %d %d 2 3 0.001\n" % (f + 1, n))
            for r in range(3):
                fh.write(" ".join(fmts[(f + r) % len(fmts)] % v for v in lat[f, r]) + "\n")
            for a in range(n):
                fh.write("C%d   %d  12.011  0.1\n" % (a % 7, a + 1))
                fh.write("  ".join(fmts[(f + a) % len(fmts)] % v for v in X[f, a]) + "\n")
                fh.write("1.0 2.0 3.0\n4.0 5.0 6.0\n")
    t = DLPOLY(path)
    assert t.no_of_frames == F and t._frame_keytrj == 2 and t._frame_imcon == 3
    xyz, lattice = t._decode_cells(list(range(F)))
    for f in range(F):
        fr = t._decode_frame(t._frame_bytes(f))                 # Python reference decode (float())
        np.testing.assert_array_equal(xyz[f], fr["coordinates"])
        np.testing.assert_array_equal(lattice[f], fr["lattice"])
        assert fr["velocities"].shape == (n, 3) and fr["forces"].shape == (n, 3)
    np.testing.assert_array_equal(t._decode_coordinates([3, 0]), xyz[[3, 0]])


def test_fixed_width_template_decoder_and_its_fallbacks(tmp_path):
    """DL_POLY's own %12.4E layout takes the template path of pwb_history_decode (tokens read at the END offsets of
    frame 0, frame shape checked byte by byte).  Signs, '+' signs and exponents differ from frame 0; frames damaged
    WITHOUT changing their length (a moved line break, a letter for a digit, ',' for the exponent sign, '!' for the
    sign, a tab) must not be mis-read: they go to the generic parser, which decodes them like float() or reports them."""
    from pywindow_b200.trajectory import DLPOLY, _TrajectoryError
    rng = np.random.default_rng(5)
    n, F = 23, 40
    X = rng.normal(0, 8, size=(F, n, 3)) * 10.0 ** rng.integers(-6, 4, size=(F, n, 1))
    X[0] = np.abs(X[0])                                        # the template frame has no minus sign at all
    X[3, 2] = 0.0

    def frame_text(f, x):
        out = ["timestep %9d %9d %9d %9d %11.6f\n" % (f + 1, n, 0, 0, 0.001)]
        for a in range(n):
            out.append("%-8s%10d%12.6f%12.6f\n" % ("C%d" % (a % 3), a + 1, 12.011, 0.0))
            out.append("%12.4E%12.4E%12.4E\n" % tuple(x[a]))
        return "".join(out)

    frames = [frame_text(f, X[f]) for f in range(F)]
    frames[5] = frames[5].replace(" 0.", "+0.", 1) if " 0." in frames[5] else frames[5]
    lens = {len(t) for t in frames}
    assert len(lens) == 1
    head = "title\n%10d%10d%10d\n" % (0, 0, n)

    def decode(fr):
        path = tmp_path / "HISTORY"
        path.write_text(head + "".join(fr))
        t = DLPOLY(path)
        ref = [t._decode_frame(t._frame_bytes(f))["coordinates"] for f in range(len(fr))] if fr is frames else None
        return t._decode_coordinates(list(range(len(fr)))), ref

    xyz, ref = decode(frames)
    for f in range(F):
        np.testing.assert_array_equal(xyz[f], ref[f])           # bit-identical to float(), -0.0 included
    assert (np.signbit(xyz) == np.signbit(np.array(ref))).all()

    def damaged(f, old, new, which=3):
        """same-length replacement inside coordinate line `which` of frame f"""
        lines = frames[f].split("\n")
        k = 2 * which + 2
        assert old in lines[k] and len(old) == len(new)
        lines[k] = lines[k].replace(old, new, 1)
        fr = list(frames)
        fr[f] = "\n".join(lines)
        return fr

    for old, new in (("E+", "E,"), ("E-", "E,"), ("E", "F"), (".", ","), (" ", "!"), (" ", "\t")):
        f = next(f for f in range(1, F) if old in frames[f].split("\n")[8] and (old != " " or frames[f].split("\n")[8][1] == " "))
        fr = damaged(f, old, new) if old != " " else None
        if fr is None:                                          # the sign column of the first token
            lines = frames[f].split("\n"); lines[8] = " " + new + lines[8][2:]; fr = list(frames); fr[f] = "\n".join(lines)
        if new == "\t":                                         # a tab is white space to float(): still decodes, equal
            got, _ = decode(fr)
            np.testing.assert_array_equal(got, xyz)
        else:
            with pytest.raises(_TrajectoryError):
                decode(fr)
    # a line break moved inside a header line: same length, same number of lines
    fr = list(frames)
    lines = fr[7].split("\n")
    lines[3] = lines[3][:5] + "\n" + lines[3][5:-1]            # breaks the header in two, drops its last character
    fr[7] = "\n".join(lines)
    assert len(fr[7]) == len(frames[7])
    path = tmp_path / "HISTORY"
    path.write_text(head + "".join(fr))
    t = DLPOLY(path)
    try:
        want = np.asarray(t._decode_frame(t._frame_bytes(7))["coordinates"], dtype=np.float64)   # line-based, like the reference
    except Exception:
        want = None
    if want is None or want.shape != (n, 3):
        with pytest.raises(_TrajectoryError):
            t._decode_coordinates([0, 7])
    else:
        np.testing.assert_array_equal(t._decode_coordinates([0, 7])[1], want)
        assert not np.array_equal(want, xyz[7])                 # (the shifted columns were not read silently)


def test_direction_bin_formula_matches_atan2():
    """csrc/pwb_device.cuh:dg_bin locates the phi bin of a direction with three comparisons against tan(k pi/16)
    per quadrant instead of atan2; the restatement here must agree with floor((atan2(y, x) + pi) / (pi/16))."""
    rng = np.random.default_rng(7)
    v = rng.normal(size=(500000, 2))
    x, y = v[:, 0], v[:, 1]
    ref = np.clip(np.floor((np.arctan2(y, x) + np.pi) * (32 / (2 * np.pi))).astype(int), 0, 31)
    a, c = np.abs(x), np.abs(y)
    t1, t2, t3 = 0.19891236737965800, 0.41421356237309503, 0.66817863791929890
    q8 = np.where(c <= a, (c > a * t1).astype(int) + (c > a * t2) + (c > a * t3),
                  7 - ((a > c * t1).astype(int) + (a > c * t2) + (a > c * t3)))
    ip = np.where(y >= 0, np.where(x >= 0, 16 + q8, 31 - q8), np.where(x >= 0, 15 - q8, q8))
    assert (ip == ref).all()


def test_native_xyz_and_pdb_decoders_are_bit_exact(tmp_path):
    """pwb_xyz_index / pwb_xyz_decode / pwb_pdb_decode (C++, threaded) against the Python per-frame decoders:
    awkward number formats, blank lines between frames, a numeric comment line (the reference's XYZ mapper
    mis-splits there), mixed atom counts (Python fallback), malformed frames."""
    from pywindow_b200.trajectory import PDB, XYZ, Trajectory, _FormatError, _TrajectoryError
    rng = np.random.default_rng(5)
    n, F = 23, 9
    fmts = ["%.17E", "%12.4E", "%20.10f", "%.3f", "%.17g", "%25.16e"]
    X = rng.normal(0, 30, size=(F, n, 3)) * 10.0 ** rng.integers(-8, 4, size=(F, n, 1))
    path = tmp_path / "t.xyz"
    with open(path, "w") as fh:
        for f in range(F):
            fh.write("  %d  \n%s\n" % (n, "42" if f == 2 else "frame %d energy -1.5" % f))
            for a in range(n):
                fh.write("C%d\t" % (a % 3) + "  ".join(fmts[(f + a) % len(fmts)] % v for v in X[f, a]) + "  extra\n")
            if f % 3 == 0:
                fh.write("\n  \n")
    t = XYZ(path)
    assert t.no_of_frames == F and (t._natoms == n).all()
    native = t._decode_coordinates(list(range(F)))
    np.testing.assert_array_equal(native, Trajectory._decode_coordinates(t, list(range(F))))   # Python float()
    np.testing.assert_array_equal(t._decode_coordinates([4, 1]), native[[4, 1]])
    # mixed atom counts: per-frame map is still right, batch decode refuses / falls back consistently
    with open(tmp_path / "mixed.xyz", "w") as fh:
        fh.write("1\nc\nC 0 0 0\n2\nc\nC 0 0 0\nH 1 0 0\n")
    m = XYZ(tmp_path / "mixed.xyz")
    assert m.no_of_frames == 2 and list(m._natoms) == [1, 2]
    assert m._decode_coordinates([1]).shape == (1, 2, 3)
    with pytest.raises(_TrajectoryError):
        m._decode_coordinates([0, 1])
    (tmp_path / "bad.xyz").write_text("3\nc\nC 0 0 0\nC 1 0 0\nC x 0 0\n")
    with pytest.raises(_TrajectoryError):
        XYZ(tmp_path / "bad.xyz")._decode_coordinates([0])
    (tmp_path / "bad2.xyz").write_text("abc\nc\n")
    with pytest.raises(_FormatError):
        XYZ(tmp_path / "bad2.xyz")
    # PDB: fixed columns, fields that touch each other, HETATM records, other records in between
    ppath = tmp_path / "t.pdb"
    Xp = np.round(rng.uniform(-999, 9999, size=(F, n, 3)), 3)
    with open(ppath, "w") as fh:
        for f in range(F):
            fh.write("REMARK frame %d\nCRYST1   30.000   30.000   30.000  90.00  90.00  90.00 P 1           1\n" % f)
            for a in range(n):
                rec = "HETATM" if a % 4 == 0 else "ATOM  "
                fh.write("%s%5d  C   MOL A   1    %8.3f%8.3f%8.3f  1.00  0.00           C\n" % ((rec, a + 1) + tuple(Xp[f, a])))
            fh.write("TER\nEND\n")
    p = PDB(ppath)
    assert p.no_of_frames == F
    native = p._decode_coordinates(list(range(F)))
    np.testing.assert_array_equal(native, Trajectory._decode_coordinates(p, list(range(F))))
    np.testing.assert_array_equal(native, Xp)


def test_pdb_model_endmdl_files_map_to_whole_frames(md):
    """Multi-model PDB with MODEL/ENDMDL pairs and a closing END: frames are cut on whole-line END /
    ENDMDL records only, the trailing END does not become an atom-less extra frame, and every frame decodes."""
    tmp, el, X, syn = md
    from pywindow_b200.trajectory import PDB
    syn.write_pdb_trajectory(tmp / "plain.pdb", el, X)
    frames = [f for f in re.split(r"^END[ \t]*\n", (tmp / "plain.pdb").read_text(), flags=re.M) if f.strip()]
    assert len(frames) == 3
    text = "".join(f"MODEL     {k + 1:4d}\n{body}ENDMDL\n" for k, body in enumerate(frames)) + "END\n"
    (tmp / "models.pdb").write_text(text)
    t = PDB(tmp / "models.pdb")
    assert t.no_of_frames == 3 and sorted(k for k in t.trajectory_map) == [0, 1, 2]
    got = t._decode_coordinates([0, 1, 2])
    np.testing.assert_allclose(got, X, atol=6e-4)
    # a residue / atom name that merely starts with END must not split a frame
    weird = text.replace("MODEL        2\n", "MODEL        2\nREMARK ENDMDL is not a record here\n")
    (tmp / "weird.pdb").write_text(weird)
    assert PDB(tmp / "weird.pdb").no_of_frames == 3


def test_lazyframe_len_and_override_semantics():
    """LazyFrame counts a key once whether it is pending, loaded or both; analysis(override=True) replaces the
    frame entry instead of merging into it (trajectory.py:483)."""
    from pywindow_b200 import _lib
    from pywindow_b200.trajectory import LazyFrame, Trajectory
    rec = np.zeros((3, _lib.REC_WIN0 + 4 * 16))
    rec[:, _lib.REC_STATUS] = _lib.ST_NO_CLEAR_RAY
    rec[:, _lib.REC_NATOMS] = [5, 6, 7]
    fr = LazyFrame()
    fr._add(0, rec, 0)
    fr._add(1, rec, 1)
    assert len(fr) == 2
    assert fr[1]["no_of_atoms"] == 6 and len(fr) == 2
    fr._add(1, rec, 2)                       # same key pending again: still two keys
    assert len(fr) == 2 and fr[1]["no_of_atoms"] == 7

    class Fake(Trajectory):
        def __init__(self):
            self.analysis_output, self.no_of_frames, self.calls = {}, 2, 0

        def _analyse_modular(self, frames, *a):
            return rec[:2 * len(frames)], [(f, m) for f in frames for m in (0, 1)]

        def _analyse_single(self, frames, *a):
            return rec[:len(frames)], [(f, "0") for f in frames]

    t = Fake()
    t.analysis(frames=[0], modular=True)
    assert sorted(t.analysis_output[0].keys()) == [0, 1]
    t.analysis(frames=[0], modular=False, override=True)
    assert list(t.analysis_output[0].keys()) == ["0"]          # no stale modular keys
    t.analysis(frames=[0, 1], modular=False)                   # override=False keeps frame 0, adds frame 1
    assert list(t.analysis_output[1].keys()) == ["0"]


def test_analysis_output_is_a_lazy_dict():
    """AnalysisOutput parks whole record tables and behaves like the reference's plain
    {frame: {molecule: properties}} dict on first access (len / in / iteration / dict() / json)."""
    import json
    from pywindow_b200 import _lib, utilities as U
    from pywindow_b200.trajectory import AnalysisOutput
    rec = np.zeros((4, _lib.REC_WIN0 + 4 * 16))
    rec[:, _lib.REC_STATUS] = _lib.ST_NO_CLEAR_RAY
    rec[:, _lib.REC_NATOMS] = [3, 4, 5, 6]
    out = AnalysisOutput()
    assert not out and len(out) == 0
    out._add_block(rec[:2], [(7, "0"), (9, "0")])
    assert bool(out)
    tab, keys = out.record_table()
    assert tab.shape == (2, rec.shape[1]) and keys == [(7, "0"), (9, "0")]
    assert 7 in out and 8 not in out and len(out) == 2 and sorted(out) == [7, 9]
    out._add_block(rec[2:], [(9, 0), (9, 1)])                    # frame 9 re-analysed (modular): replaced
    assert sorted(out[9].keys()) == [0, 1] and out[9][1]["no_of_atoms"] == 6
    plain = dict(out)
    assert sorted(plain) == [7, 9] and plain[7]["0"]["no_of_atoms"] == 3
    out[11] = {"0": {"no_of_atoms": 1}}
    assert out.get(11)["0"]["no_of_atoms"] == 1 and out.get(12) is None
    text = json.dumps({str(f): {str(k): v for k, v in dict(m.items()).items()} for f, m in out.items()}, default=U.to_list)
    assert json.loads(text)["7"]["0"]["windows"]["diameters"] is None


def test_writers_reproduce_the_reference_files(tmp_path):
    """Output.dump2file against files written by the reference's Output for the same system (tests/golden/writer_ref.*,
    oracle/make_golden.py --only writers): byte identical (io_tools.py:349-493)."""
    from pywindow_b200.io_tools import Output
    g = load_golden("md_frame00")
    keys = np.array([e + str(i % 7) for i, e in enumerate(g["elements"])])
    system = {"elements": g["elements"].copy(), "atom_ids": keys, "coordinates": g["coords"] * 3.7 - 11.0,
              "unit_cell": np.array([24.8, 25.1, 26.0, 90.0, 95.5, 101.25])}
    gold = os.path.join(ROOT, "tests", "golden")
    Output().dump2file(system, tmp_path / "w.pdb", atom_ids_key="atom_ids", remarks=["first", "second"], space_group="P21")
    assert (tmp_path / "w.pdb").read_bytes() == open(os.path.join(gold, "writer_ref.pdb"), "rb").read()
    Output().dump2file(system, tmp_path / "w.xyz", atom_ids_key="elements", remarks=["a", "b"])
    assert (tmp_path / "w.xyz").read_bytes() == open(os.path.join(gold, "writer_ref.xyz"), "rb").read()
