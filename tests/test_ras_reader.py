""""Next" rows N3 / N4: this repo's h5py-free HEC-RAS reader (io/hdf5_mini.py, io/ras.py) and its
vectorised IC / BC ingestion, against what the UNMODIFIED reference read from the same files (the raw arrays
and input arrays stored in tests/golden by tools/make_golden.py).  Needs the reference's fixture files, so it
runs in the build container only (skipped where /root/reference is absent, e.g. on the GPU box)."""
from pathlib import Path

import numpy as np
import pytest

from tests.helpers import load_golden

DATA = Path("/root/reference/tests/data/simple_test_cases")
CASES = {   # golden case -> (plan dir, tag, datetime_range)
    "p02_uniform100": ("plan02_2x1", "p02", None),
    "p01_uniform100": ("plan01_10x5", "p01", (0, 300)),
    "p03_uniform100": ("plan03_2x1", "p03", (0, 400)),
}

pytestmark = pytest.mark.skipif(not DATA.is_dir(), reason="reference fixtures not present")


@pytest.mark.parametrize("case", sorted(CASES))
def test_reader_and_ingestion_match_what_the_reference_read(case):
    from clearwater_riverine_b200.io import ras
    d, tag, dtr = CASES[case]
    base = DATA / d
    g = load_golden(case)
    plan = ras.read_ras_plan(base / f"clearWaterTestCases.{tag}.hdf", dtr)
    assert np.array_equal(plan.f1, g["f1"]) and np.array_equal(plan.f2, g["f2"])
    assert np.array_equal(plan.face_x, g["face_x"]) and np.array_equal(plan.face_y, g["face_y"])
    assert np.array_equal(plan.time_seconds, g["time_seconds"])
    for mine, ref in ((plan.face_flow, g["face_flow"]), (plan.edge_velocity, g["edge_velocity"]), (plan.volume, g["volume"])):
        assert mine.dtype == np.float32 and np.array_equal(mine, ref)
    assert plan.nreal + 1 == int(g["f1"].max()) + 1
    # boundary lines: the same faces the reference attached its BC series to
    bf = plan.boundary_faces()
    for name, face in zip(g["bc_names"], g["bc_faces"]):
        assert int(face) in set(int(x) for x in bf[str(name)])
    # IC / BC ingestion -> the (T,F) input_array of the constituent, bit for bit
    cname = str(g["constituents"][0])
    inp = ras.build_input_array(len(plan.time), len(plan.face_x), plan.time, plan.f2, plan.boundary_data,
                                base / f"cwr_initial_conditions_{tag}.csv", base / f"cwr_boundary_conditions_{tag}.csv")
    assert np.array_equal(inp, g[f"input_{cname}"], equal_nan=True)


def test_missing_file_raises_like_the_reference():
    from clearwater_riverine_b200.io import ras
    with pytest.raises(FileNotFoundError):        # reference io/inputs.py:39-44
        ras.read_ras_plan(DATA / "does_not_exist.hdf")
