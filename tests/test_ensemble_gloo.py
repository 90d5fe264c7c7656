"""N > 1 host logic on CPU: two gloo ranks shard the independent units (constituents / scenarios),
compute their mass-balance rows with the oracle (the GPU is not available here) and reduce them the
way bench.py does over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from clearwater_riverine_b200 import ensemble


def test_shard_units_partitions_exactly():
    for n_units, world in [(64, 8), (64, 3), (5, 8), (16, 1), (0, 4)]:
        seen = []
        for r in range(world):
            block = ensemble.shard_units(n_units, world, r)
            seen.extend(block)
            for u in block:
                assert ensemble.owner_of(u, n_units, world) == r
        assert seen == list(range(n_units))
        sizes = [len(ensemble.shard_units(n_units, world, r)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    assert ensemble.ensemble_plan(10, 4) == [(0, 3), (3, 3), (6, 2), (8, 2)]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_units, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from clearwater_riverine_b200 import synthetic
        from oracle import reference_step as ref
        plan = synthetic.make_plan(12, 8, 5, seed=4)
        D = 0.1
        adv, _, _, cdiff, dt = ref.derive_coefficients(plan.face_flow, plan.edge_velocity, plan.face_x, plan.face_y,
                                                       plan.f1, plan.f2, D, plan.time_seconds)
        mesh = ref.HydroMesh(plan.f1, plan.f2, plan.n_face, adv, cdiff, plan.edge_velocity, plan.volume, dt, D)
        base = synthetic.make_inputs(plan, 1, seed=4)[0]
        scales = np.exp(np.random.default_rng(7).normal(0.0, 0.5, size=n_units))      # same on every rank
        mine = ensemble.shard_units(n_units, world, rank)
        inputs = ensemble.scenario_inputs(base, plan.n_real, scales[mine.start:mine.stop])
        local = {}
        n = plan.n_real
        for unit, inp in zip(mine, inputs):
            model = ref.OracleRiverine(mesh, {"c": inp})
            for _ in range(plan.n_time - 1):
                model.update()
            c = model.constituent_dict["c"].concentration
            local[unit] = (float((mesh.vol[0][:n] * c[0][:n]).sum()), float((mesh.vol[-1][:n] * c[-1][:n]).sum()))
        table = ensemble.reduce_mass_balance(local, n_units)
        slowest = ensemble.max_over_ranks(1.0 + rank)
        np.save(os.path.join(out_dir, f"table_{rank}.npy"), table)
        np.save(os.path.join(out_dir, f"slow_{rank}.npy"), np.array([slowest]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_ensemble_reduction(tmp_path):
    world, n_units = 2, 5
    mp.start_processes(_worker, args=(world, _free_port(), n_units, str(tmp_path)), nprocs=world, join=True,
                       start_method="spawn")
    t0, t1 = np.load(tmp_path / "table_0.npy"), np.load(tmp_path / "table_1.npy")
    assert t0.shape == (n_units, 2)
    assert np.array_equal(t0, t1)                       # every rank ends up with the full table
    assert np.all(t0[:, 0] > 0) and np.all(t0[:, 1] > 0)
    assert np.load(tmp_path / "slow_0.npy")[0] == 2.0   # max over ranks
    # the same initial condition in every scenario -> identical starting mass; different BCs -> different end mass
    assert np.allclose(t0[:, 0], t0[0, 0], rtol=0, atol=0)
    assert len(np.unique(np.round(t0[:, 1], 6))) == n_units
