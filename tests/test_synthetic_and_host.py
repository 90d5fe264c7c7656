"""CPU checks of the synthetic HEC-RAS-like generator (the benchmark configurations run on its meshes) and of the
host mirror's argument handling (no device needed for either)."""
import numpy as np
import pytest

from clearwater_riverine_b200 import synthetic


def test_generator_follows_the_reference_conventions():
    plan = synthetic.make_plan(24, 17, 6, tri_fraction=0.2, dry_fraction=0.03, seed=8)
    n, F, E = plan.n_real, plan.n_face, plan.n_edge
    assert plan.f1.dtype == np.int32 and plan.f2.dtype == np.int32
    assert plan.f1.max() == n - 1 and plan.f1.min() >= 0            # nreal := max(edges_face1)  (io/hdf.py:268-269)
    ghost = plan.f2 >= n
    assert np.array_equal(np.sort(plan.f2[ghost]), np.arange(n, F))     # one ghost cell per perimeter edge
    assert plan.face_flow.dtype == np.float32 and plan.face_flow.shape == (6, E)
    assert plan.volume.shape == (6, F) and plan.edge_velocity.shape == (6, E)
    assert np.all(np.sign(plan.edge_velocity) == np.sign(plan.face_flow))   # the fixtures' property
    # dry cells: zero volume and no flow across any of their faces
    dry = np.zeros(F, bool); dry[plan.dry_cells] = True
    assert len(plan.dry_cells) > 0 and np.all(plan.volume[:, plan.dry_cells] == 0)
    touches_dry = dry[plan.f1] | dry[np.minimum(plan.f2, F - 1)] & (plan.f2 < n)
    assert np.all(plan.face_flow[:, touches_dry] == 0)


def test_generator_continuity_to_float32_rounding():
    """V[t+1] = V[t] - dt * (net outflow at t): what keeps a uniform concentration uniform under steady flow."""
    plan = synthetic.make_plan(20, 15, 8, tri_fraction=0.1, dry_fraction=0.0, seed=5)
    n = plan.n_real
    dt = np.diff(plan.time_seconds)
    for t in range(plan.n_time - 1):
        q = plan.face_flow[t].astype(np.float64)
        out = np.bincount(plan.f1, q, n)                                  # flow > 0 leaves f1
        inn = np.bincount(plan.f2[plan.f2 < n], q[plan.f2 < n], n)
        v_pred = plan.volume[t, :n].astype(np.float64) - dt[t] * (out - inn)
        err = np.abs(v_pred - plan.volume[t + 1, :n]).max() / plan.volume[:, :n].max()
        assert err < 5e-6, (t, err)


def test_benchmark_presets_have_the_named_sizes():
    ohio = synthetic.ohio_like(4)
    assert ohio.n_real == 2943                                            # examples/Ohio River.ipynb cell 9
    import bench
    plan, K = bench.workload_plan("1m16", 3, seed=2, scale=0.05)          # scaled down: shape of the call only
    assert K == 16 and plan.n_real > 2000
    assert bench.workload_plan("ens64", 3, seed=2)[1] == 64


def test_make_inputs_keep_zero_for_unset():
    plan = synthetic.make_plan(10, 8, 5, seed=1)
    inp = synthetic.make_inputs(plan, 3, seed=1)
    n = plan.n_real
    assert inp.shape == (3, 5, plan.n_face)
    assert np.all(inp[:, 0, :n] > 0)                      # IC strictly positive: zero means "not set" (linalg.py:199-200)
    assert np.all(inp[:, 1:, :n] == 0)                    # nothing on real cells after row 0
    bc_cols = np.concatenate([plan.f2[plan.boundary_faces[k]] for k in ("upstream", "downstream")])
    assert np.all(inp[:, :, bc_cols] > 0)
    other = np.setdiff1d(np.arange(n, plan.n_face), bc_cols)
    assert np.all(inp[:, :, other] == 0)


def test_host_mirror_rejects_bad_arguments_like_the_reference():
    from clearwater_riverine_b200 import ClearwaterRiverine
    with pytest.raises(TypeError):                        # transport.py:121-123
        ClearwaterRiverine()
    with pytest.raises(NotImplementedError):
        ClearwaterRiverine(mesh_file_path="saved.zarr")
    with pytest.raises(FileNotFoundError):                # io/inputs.py:39-44
        ClearwaterRiverine(flow_field_file_path="/nonexistent/plan.hdf", diffusion_coefficient_input=0.1,
                           constituent_dict={"c": {"initial_conditions": "ic.csv", "boundary_conditions": "bc.csv"}})
