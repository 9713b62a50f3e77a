"""Host-side mirror of the reference interface checked against the LIVE reference (only in the build
container, where /root/reference exists; skipped on the GPU box).  No GPU compute."""
import copy
import os

import numpy as np
import pytest

from conftest import load_golden

refharness = pytest.importorskip("refharness")
if not refharness.reference_available():
    pytest.skip("reference not present", allow_module_level=True)

INP = os.path.join(refharness.REF_ROOT, "examples", "data", "input")


@pytest.fixture(scope="module")
def ref():
    return refharness.load_reference()


def test_supercell_and_cell_algebra(ref):
    pw, U = ref
    import pywindow_b200.trajectory as T
    from pywindow_b200 import periodic, utilities as V
    g = load_golden("per_tatver_cell")
    sysd = {"elements": g["elements"], "atom_ids": g["elements"], "coordinates": g["coords"], "unit_cell": g["unit_cell"]}
    a = U.create_supercell(copy.deepcopy(sysd))
    b = periodic.create_supercell(copy.deepcopy(sysd))
    np.testing.assert_allclose(b["coordinates"], a["coordinates"], atol=1e-10)
    assert (b["elements"] == a["elements"]).all()
    np.testing.assert_allclose(b["lattice"], a["lattice"], atol=1e-12)
    np.testing.assert_allclose(V.unit_cell_to_lattice_array(g["unit_cell"]), U.unit_cell_to_lattice_array(g["unit_cell"]))
    np.testing.assert_allclose(V.lattice_array_to_unit_cell(a["lattice"]), U.lattice_array_to_unit_cell(a["lattice"]))
    s1 = pw.make_supercell(copy.deepcopy(sysd), [2, 1, 1]).system
    s2 = T.make_supercell(copy.deepcopy(sysd), [2, 1, 1]).system
    np.testing.assert_allclose(s2["coordinates"], s1["coordinates"], atol=1e-10)


def test_history_decode_and_deciphering(ref):
    pw, U = ref
    import pywindow_b200 as mine
    path = os.path.join(INP, "HISTORY_singlemol_short")
    a, b = pw.DLPOLY(path), mine.DLPOLY(path)
    assert a.no_of_frames == b.no_of_frames == 20 and a.no_of_atoms == b.no_of_atoms
    for f in (0, 7, 19):
        ra = a.get_frames(f, swap_atoms={"he": "H"}, forcefield="opls")[f].system
        rb = b.get_frames(f, swap_atoms={"he": "H"}, forcefield="opls")[f].system
        np.testing.assert_array_equal(rb["coordinates"], ra["coordinates"])
        assert list(rb["elements"]) == list(ra["elements"]) and list(rb["atom_ids"]) == list(ra["atom_ids"])
    np.testing.assert_array_equal(b._decode_coordinates([3, 4]), np.stack(
        [a.get_frames(k)[k].system["coordinates"] for k in (3, 4)]))
    for key, ff in (("C1", "DLF"), ("Cl12", "DL_F"), ("ca", "opls"), ("ni", "OPLSAA"), ("HC", "OPLS3")):
        assert mine.utilities.decipher_atom_key(key, ff) == U.decipher_atom_key(key, ff)


def test_file_readers(ref):
    pw, U = ref
    import pywindow_b200 as mine
    for fn in ("PUDXES.xyz", "MIBQAR.pdb", "EPIRUR_no_solvent.pdb"):
        a = pw.MolecularSystem.load_file(os.path.join(INP, fn)).system
        b = mine.MolecularSystem.load_file(os.path.join(INP, fn)).system
        np.testing.assert_array_equal(b["coordinates"], a["coordinates"])
        assert list(b["elements"]) == list(a["elements"])
        if "unit_cell" in a:
            np.testing.assert_array_equal(b["unit_cell"], a["unit_cell"])


def test_tables_are_the_reference_tables(ref):
    from pywindow._internal import tables as RT
    from pywindow_b200 import tables as MT
    assert MT.atomic_mass == RT.atomic_mass and MT.atomic_vdw_radius == RT.atomic_vdw_radius
    assert MT.atomic_covalent_radius == RT.atomic_covalent_radius
    assert MT.opls_atom_keys == RT.opls_atom_keys and MT.periodic_table == RT.periodic_table


def test_xyz_trajectory_decode_against_reference(ref, tmp_path):
    """The native XYZ frame map / decoder against the reference's own XYZ trajectory reader
    (trajectory.py:792-915) on a file it maps correctly (non-numeric comment lines)."""
    pw, U = ref
    import pywindow_b200 as mine
    from pywindow_b200 import synthetic
    gs = [load_golden(f"md_frame{i:02d}") for i in range(3)]
    X = np.stack([g["coords"] for g in gs])
    synthetic.write_xyz_trajectory(tmp_path / "t.xyz", gs[0]["elements"], X)
    a, b = pw.XYZ(str(tmp_path / "t.xyz")), mine.XYZ(tmp_path / "t.xyz")
    assert a.no_of_frames == b.no_of_frames == 3
    native = b._decode_coordinates([0, 1, 2])
    for f in range(3):
        ra = a.get_frames(f)[f].system
        np.testing.assert_array_equal(native[f], np.asarray(ra["coordinates"], dtype=np.float64))
        rb = b.get_frames(f)[f].system
        assert list(rb["atom_ids"]) == list(ra["atom_ids"])      # the reference keeps XYZ symbols as atom_ids


def test_writers_are_byte_identical(ref, tmp_path):
    """Output._save_pdb / _save_xyz / dump2file and Molecule.dump_molecule(include_coms=True) against the
    reference's writers (io_tools.py:267-493, molecular.py:419-545): same bytes for the same system."""
    pw, U = ref
    import pywindow_b200 as mine
    from pywindow._internal.io_tools import Output as RefOutput
    from pywindow_b200.io_tools import Output
    g = load_golden("md_frame00")
    rng = np.random.default_rng(4)
    keys = np.array([e + str(i % 7) for i, e in enumerate(g["elements"])])
    system = {"elements": g["elements"].copy(), "atom_ids": keys, "coordinates": g["coords"] * 3.7 - 11.0,
              "unit_cell": np.array([24.8, 25.1, 26.0, 90.0, 95.5, 101.25])}
    cases = [dict(atom_ids_key="elements"), dict(atom_ids_key="atom_ids", remarks=["first", "second"], space_group="P21"),
             dict(atom_ids_key="atom_ids", remarks="one line", resname="CC3", chainid="B", resseq=12, cryst="nothing"),
             dict(atom_ids_key="elements", elements_key="atom_ids", decipher=True, forcefield="DLF", remarks=3.5)]
    for k, kw in enumerate(cases):
        a, b = tmp_path / f"ref{k}.pdb", tmp_path / f"mine{k}.pdb"
        RefOutput().dump2file(copy.deepcopy(system), a, **kw)
        Output().dump2file(copy.deepcopy(system), b, **kw)
        assert a.read_bytes() == b.read_bytes(), kw
    for k, kw in enumerate([dict(), dict(remarks=["a", "b"]), dict(remarks="r", elements_key="atom_ids", decipher=True, forcefield="DLF")]):
        a, b = tmp_path / f"ref{k}.xyz", tmp_path / f"mine{k}.xyz"
        RefOutput().dump2file(copy.deepcopy(system), a, atom_ids_key="elements", **kw)
        Output().dump2file(copy.deepcopy(system), b, atom_ids_key="elements", **kw)
        assert a.read_bytes() == b.read_bytes(), kw
    with pytest.raises(FileExistsError):
        Output().dump2file(system, tmp_path / "mine0.pdb", atom_ids_key="elements")
    # overlay of COM (He), optimised pore centre (Ne) and window centres (Ar1..) -- properties injected, no GPU
    props = {"centre_of_mass": rng.normal(size=3), "pore_diameter_opt": {"centre_of_mass": rng.normal(size=3)},
             "windows": {"centre_of_mass": rng.normal(size=(4, 3)) * 5}}
    for with_ids in (True, False):
        mol = {"elements": system["elements"].copy(), "coordinates": system["coordinates"].copy()}
        if with_ids:
            mol["atom_ids"] = keys.copy()
        ra = pw.Molecule(copy.deepcopy(mol), "sys", 0)
        rb = mine.Molecule(copy.deepcopy(mol), "sys", 0)
        ra.properties.update(copy.deepcopy(props))
        rb.properties.update(copy.deepcopy(props))
        a, b = tmp_path / f"ref_coms{with_ids}.pdb", tmp_path / f"mine_coms{with_ids}.pdb"
        ra.dump_molecule(a, include_coms=True)
        rb.dump_molecule(b, include_coms=True)
        assert a.read_bytes() == b.read_bytes()
        a2, b2 = tmp_path / f"ref_plain{with_ids}.pdb", tmp_path / f"mine_plain{with_ids}.pdb"
        ra.dump_molecule(a2)
        rb.dump_molecule(b2)
        assert a2.read_bytes() == b2.read_bytes()
