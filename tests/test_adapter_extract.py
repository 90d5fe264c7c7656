"""attach(model): the arrays read from a LIVE, unmodified reference object (built here through oracle/refshim from
/root/reference; skipped where the reference tree is absent, e.g. on the GPU box) equal the golden vectors the reference
produced -- i.e. the adapter hands the device exactly what the reference's own linalg would read."""
import contextlib
import io
import warnings
from pathlib import Path

import numpy as np
import pytest

from tests.helpers import load_golden

REF = Path("/root/reference/src/clearwater_riverine/transport.py")
DATA = Path("/root/reference/tests/data/simple_test_cases")


@pytest.mark.skipif(not REF.is_file(), reason="needs the reference tree (build container only)")
def test_extract_model_arrays_from_the_real_reference_object():
    from clearwater_riverine_b200 import extract_model_arrays
    from oracle.refshim import load_reference
    cwr = load_reference()
    base = DATA / "plan02_2x1"
    cdict = {"tracer": {"initial_conditions": str(base / "cwr_initial_conditions_p02.csv"),
                        "boundary_conditions": str(base / "cwr_boundary_conditions_p02.csv"), "units": "mg/L"}}
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = cwr.ClearwaterRiverine(flow_field_file_path=str(base / "clearWaterTestCases.p02.hdf"),
                                       diffusion_coefficient_input=0.01, constituent_dict=cdict)
    a = extract_model_arrays(model)
    g = load_golden("p02_uniform100")
    assert a["n_real"] == 2 and a["n_face"] == 8 and a["n_time"] == 25 and a["time_step"] == 0
    assert a["constituents"] == ["tracer"] and a["diffusion_coefficient"] == float(g["diffusion_coefficient"])
    for mine, theirs in (("f1", "f1"), ("f2", "f2"), ("adv", "adv"), ("cdiff", "cdiff"), ("vel", "edge_velocity"), ("vol", "volume"),
                         ("dt", "dt")):
        assert np.array_equal(a[mine], g[theirs], equal_nan=True), mine
    assert np.array_equal(a["inputs"]["tracer"], g["input_tracer"])
