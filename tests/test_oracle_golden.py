"""Pin the CPU oracle (oracle/reference_step.py) against golden vectors produced by the
UNMODIFIED reference (tools/make_golden.py ran /root/reference's own ClearwaterRiverine).

Bit-exact: the oracle builds the same COO triplets in the same order and hands them
to the same scipy csr_matrix / spsolve, so every float must be identical.
"""
import numpy as np
import pytest

from oracle import reference_step as ref
from tests.helpers import GOLDEN_CASES, golden_mesh, golden_overrides, load_golden, same


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_derived_coefficients_match_reference(case):
    g = load_golden(case)
    adv, area, dist, cdiff, dt = ref.derive_coefficients(
        g["face_flow"], g["edge_velocity"], g["face_x"], g["face_y"], g["f1"], g["f2"],
        float(g["diffusion_coefficient"]), g["time_seconds"])
    assert adv.dtype == np.float32 and cdiff.dtype == np.float64
    assert same(adv, g["adv"])
    assert same(area, g["area"])
    assert same(dist, g["dist"])
    assert same(cdiff, g["cdiff"])
    assert same(dt, g["dt"])


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_free_running_trajectory_is_bit_identical(case):
    g = load_golden(case)
    mesh = golden_mesh(g)
    names = [str(c) for c in g["constituents"]]
    model = ref.OracleRiverine(mesh, {c: g[f"input_{c}"] for c in names})
    overrides = golden_overrides(g)
    snaps = set(int(s) for s in g["snapshot_steps"])
    for t in range(mesh.n_time - 1):
        model.update(overrides.get(t))
        if t in snaps:
            A = model.last_A.copy()
            A.sum_duplicates()
            A.sort_indices()
            assert same(A.indptr, g[f"A_indptr_{t}"]) and same(A.indices, g[f"A_indices_{t}"])
            assert same(A.data, g[f"A_data_{t}"]), f"LHS differs at step {t}"
            for c in names:
                assert same(model.constituent_dict[c].b.vals, g[f"b_{c}_{t}"]), f"RHS differs at step {t}"
    assert model.time_step == mesh.n_time - 1
    for c in names:
        con = model.constituent_dict[c]
        assert same(con.concentration, g[f"conc_{c}"]), c
        assert same(con.advection_mass_flux, g[f"advflux_{c}"])
        assert same(con.diffusion_mass_flux, g[f"diffflux_{c}"])
        assert same(con.total_mass_flux, g[f"totflux_{c}"])


def test_reference_test_riverine_intent_p02():
    """tests/test_riverine.py:89-127 of the reference (stale API, same intent)."""
    g = load_golden("p02_uniform100")
    mesh = golden_mesh(g)
    inp = g["input_tracer"]
    assert inp[0, 0] == 100 and inp[0, 4] == 100 and inp[0, 6] == 100     # test_riverine_initialize
    model = ref.OracleRiverine(mesh, {"tracer": inp})
    model.update()
    c = model.constituent_dict["tracer"].concentration
    assert c[1, 0] != 0 and model.time_step == 1                          # test_riverine_update
    assert c[1, 4] == 100 and c[1, 6] == 100


def test_uniform_100_mass_end_p02():
    """tests/test_final_mass.py:29-33: Mass_end vs the 'all cells are 100' answer.  The
    reference asserts ==, which its own arithmetic misses by 2.8e-11 relative;
    the notebook print it was derived from is Mass_end(100) = 5001.221848."""
    g = load_golden("p02_uniform100")
    mesh = golden_mesh(g)
    n, T = mesh.n, mesh.n_time
    mass_end = float((mesh.vol[T - 1][:n] * g["conc_tracer"][T - 1][:n]).sum())
    mass_100 = float((mesh.vol[T - 1][:n].astype(np.float64) * 100.0).sum())
    assert abs(mass_100 - 5001.221848) < 1e-6
    assert abs(mass_end - mass_100) / mass_100 < 1e-9
    mb = ref.mass_balance(mesh, g["conc_tracer"], g["totflux_tracer"], g["face_flow"],
                          {str(nm): np.array([f]) for nm, f in zip(g["bc_names"], g["bc_faces"])})
    assert mb["Mass_end"] == pytest.approx(mass_end, rel=1e-15)
