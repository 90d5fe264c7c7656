"""Multi-GPU host logic: sharding of independent units and the mass-balance reduction.

Constituents and boundary-condition scenarios only share the mesh and the hydrodynamics
(the reference loops over them one after another against the same LHS, transport.py:231), so they
are the natural multi-GPU axis: every rank owns a contiguous block of units on a
replica of the mesh and steps them with no data-path communication; the only collective is a sum of
the per-unit mass-balance scalars (postproc_util.py:36-59, 100-143) at report time.  One process per
GPU; `torch.distributed` supplies the process group (NCCL on GPUs, gloo in the CPU tests) -- torch
is imported lazily so that the single-GPU product does not depend on it.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np


def shard_units(n_units: int, world_size: int, rank: int) -> range:
    """Contiguous, balanced block of unit indices owned by `rank` (sizes differ by at most one)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, extra = divmod(n_units, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def owner_of(unit: int, n_units: int, world_size: int) -> int:
    base, extra = divmod(n_units, world_size)
    split = extra * (base + 1)
    return unit // (base + 1) if unit < split else extra + (unit - split) // max(base, 1)


def reduce_mass_balance(local: Dict[int, Sequence[float]], n_units: int, group=None, device=None) -> np.ndarray:
    """All-reduce (sum) of per-unit mass-balance rows.

    local: {global unit index: (mass_start, mass_end, boundary_in, boundary_out, ...)} for the units this
    rank owns.  Returns the (n_units, width) table, identical on every rank.  A unit owned by nobody stays 0.
    """
    import torch
    import torch.distributed as dist
    width = max((len(v) for v in local.values()), default=0)
    w = torch.tensor([width], dtype=torch.int64, device=device)
    if dist.is_initialized():
        dist.all_reduce(w, op=dist.ReduceOp.MAX, group=group)
    width = int(w.item())
    table = torch.zeros((n_units, width), dtype=torch.float64, device=device)
    for unit, row in local.items():
        table[unit, : len(row)] = torch.as_tensor(list(row), dtype=torch.float64)
    if dist.is_initialized():
        dist.all_reduce(table, op=dist.ReduceOp.SUM, group=group)
    return table.cpu().numpy()


def max_over_ranks(value: float, group=None, device=None) -> float:
    """Timing rule of the benchmark: a multi-GPU time is the maximum over the ranks."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def scenario_inputs(base_input: np.ndarray, n_real: int, scales: Sequence[float]) -> List[np.ndarray]:
    """Boundary-condition scenarios of one constituent: the ghost-cell (BC) columns of `input_array`
    scaled per scenario, the initial condition shared (BASELINE.json configs[3])."""
    out = []
    for s in scales:
        a = base_input.copy()
        a[:, n_real:] *= s
        out.append(a)
    return out


def ensemble_plan(n_units: int, world_size: int) -> List[Tuple[int, int]]:
    """[(first unit, count)] per rank -- what bench.py prints into its config."""
    return [(r.start, len(r)) for r in (shard_units(n_units, world_size, k) for k in range(world_size))]
