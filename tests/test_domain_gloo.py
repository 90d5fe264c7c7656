"""Host side of the domain decomposition on CPU (two gloo ranks): the IPC-handle all-gather keeps rank
order, and the merge of per-rank outputs over the ownership masks the library reports (cwr_order_cells
gives the same strips cwr_dd_layout reports on a GPU) reassembles the oracle's full result exactly."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from clearwater_riverine_b200 import synthetic
        from clearwater_riverine_b200.backend import IPC_HANDLE_BYTES, order_cells
        from clearwater_riverine_b200.domain import exchange_ipc_handles, merge_owned
        # (1) handles come back in rank order on every rank
        mine = bytes([(rank * 37 + i) % 256 for i in range(IPC_HANDLE_BYTES)])
        allh = exchange_ipc_handles(mine, rank, world)
        assert len(allh) == world * IPC_HANDLE_BYTES
        for q in range(world):
            assert allh[q * IPC_HANDLE_BYTES:(q + 1) * IPC_HANDLE_BYTES] == bytes([(q * 37 + i) % 256 for i in range(IPC_HANDLE_BYTES)])
        # (2) strips of the ordering -> ownership masks in reference numbering; merge of masked per-rank arrays
        plan = synthetic.make_plan(30, 20, 4, seed=9, tri_fraction=0.2)
        n = plan.n_real
        new_of_old, _, _, part_ptr, n_send = order_cells(plan.f1, plan.f2, plan.n_face, True, 12, plan.face_flow.mean(0), world)
        owned = (new_of_old >= part_ptr[rank]) & (new_of_old < part_ptr[rank + 1])
        full = np.random.default_rng(3).random((5, n))                 # "the answer", same on every rank
        local = np.where(owned[None, :], full, np.nan)                 # a rank only holds its strip (garbage elsewhere)
        merged = merge_owned(local, owned, axis=1)
        np.save(os.path.join(out_dir, f"merged_{rank}.npy"), merged)
        np.save(os.path.join(out_dir, f"full_{rank}.npy"), full)
        np.save(os.path.join(out_dir, f"owned_{rank}.npy"), owned)
        # partial sums (mass totals of a strip) add up
        tot = merge_owned(np.array([full[0][owned].sum(), float(owned.sum())]), np.ones(2, bool))
        np.save(os.path.join(out_dir, f"tot_{rank}.npy"), tot)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_handle_exchange_and_merge(tmp_path):
    world = 2
    mp.start_processes(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    full = np.load(tmp_path / "full_0.npy")
    o0, o1 = np.load(tmp_path / "owned_0.npy"), np.load(tmp_path / "owned_1.npy")
    assert np.all(o0 ^ o1)                                   # the strips partition the cells
    assert abs(int(o0.sum()) - int(o1.sum())) <= 1
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"merged_{r}.npy"), full)
        tot = np.load(tmp_path / f"tot_{r}.npy")
        assert tot[1] == full.shape[1] and abs(tot[0] - full[0].sum()) <= 1e-12 * full[0].sum()
