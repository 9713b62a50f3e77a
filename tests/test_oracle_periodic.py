"""Pin oracle/pw_periodic_oracle.py (discrete_molecules / create_supercell restatement) against the
reference's own outputs (tests/golden/per_*.npz) and its test-suite counts
(/root/reference/tests/test_molecular.py:4467-4553: 33 fragments, 8 x 168 atoms after rebuild)."""
import numpy as np
import pytest

from conftest import golden_names, load_golden

PER = golden_names(("per_",))


@pytest.mark.parametrize("name", PER)
@pytest.mark.parametrize("mode", ["norebuild", "rebuild"])
def test_discrete_molecules(name, mode):
    import pw_periodic_oracle as PO
    g = load_golden(name)
    lat = PO.unit_cell_to_lattice_array(g["unit_cell"])
    mols = PO.discrete_molecules(g["elements"], g["coords"], lat, rebuild=(mode == "rebuild"))
    assert len(mols) == int(g[mode + "_count"])
    assert [len(m[0]) for m in mols] == list(g[mode + "_sizes"])
    assert (np.concatenate([m[0] for m in mols]) == g[mode + "_elements"]).all()
    np.testing.assert_allclose(np.concatenate([m[1] for m in mols]), g[mode + "_coords"], atol=1e-7)


def test_reference_suite_counts():
    g = load_golden("per_pudxes_cell")
    assert int(g["norebuild_count"]) == 33 and int(g["rebuild_count"]) == 8
    assert (g["rebuild_sizes"] == 168).all()
