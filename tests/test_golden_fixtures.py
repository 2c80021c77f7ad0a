"""The golden fixtures (tests/golden/): every entry cites the reference source it was transcribed from, and the
constants the parity tests use are the fixture's values."""
import numpy as np

from golden_data import GOLDEN
from problems import ELASTIC_STRIP_EXPECTED, PATCH_EXPECTED_D


def test_every_entry_cites_its_reference_source():
    for name, entry in GOLDEN.items():
        if name.startswith("_"):
            continue
        entries = entry.values() if name == "not_reproduced" else [entry]
        for e in entries:
            assert "tests/src/" in e["source"], name


def test_shared_constants_come_from_the_fixture():
    assert ELASTIC_STRIP_EXPECTED[("svk", 1)] == (6, 1.814746879163122)
    assert ELASTIC_STRIP_EXPECTED[("neohooke", 2)] == (7, 2.1944518710582974)
    assert PATCH_EXPECTED_D.shape == (16,) and PATCH_EXPECTED_D[2] == 0.001
    assert len(GOLDEN["cantilever_eas"]["cases"]) == 4
    assert np.array(GOLDEN["cube_vertex_stress"]["values"]).shape == (8, 6)
    assert np.array(GOLDEN["square_vertex_stress"]["eas4"]).shape == (4, 3)
