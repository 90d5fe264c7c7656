#!/usr/bin/env python
"""Smallest cases that exercise every kernel family, for compute-sanitizer (one tool per run):

  compute-sanitizer --tool memcheck  python tools/sanitize_case.py
  compute-sanitizer --tool racecheck python tools/sanitize_case.py

On-chip solver (k_solve_tiny), one-CTA-per-column solver through L2 (k_solve_small), and the multi-CTA path with the
three Gauss-Seidel sweep kernels (grid barrier / neighbour flags / software-pipelined: 23 cooperating CTAs), both
solvers, the asynchronous output path and the mass-balance reductions.  Each case is checked against the oracle."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    from clearwater_riverine_b200 import TransportBackend, synthetic
    from oracle import reference_step as ref
    D = 0.1
    cases = [("tiny", (24, 16), 2, dict()), ("small", (24, 16), 2, dict(precond_sweep=0)),
             ("pipelined strips + defect correction", (120, 90), 4, dict(solver_path=1)),
             ("strips + BiCGSTAB", (120, 90), 3, dict(solver_path=1, solver=1, precond_sync=2)),
             ("grid barrier + defect correction", (120, 90), 2, dict(solver_path=1, precond_sync=1))]
    for name, (nx, ny), K, opts in cases:
        plan = synthetic.make_plan(nx, ny, 4, seed=11, dry_fraction=0.02)
        adv, _, _, cdiff, dt = ref.derive_coefficients(plan.face_flow, plan.edge_velocity, plan.face_x, plan.face_y,
                                                       plan.f1, plan.f2, D, plan.time_seconds)
        inputs = synthetic.make_inputs(plan, K, seed=11)
        be = TransportBackend(plan.f1, plan.f2, plan.n_face, plan.n_time, K, D, device=0, **opts)
        be.set_hydro(0, adv, cdiff, plan.edge_velocity, plan.volume, dt)
        for k in range(K):
            be.set_inputs(k, inputs[k])
        mesh = ref.HydroMesh(plan.f1, plan.f2, plan.n_face, adv, cdiff, plan.edge_velocity, plan.volume, dt, D)
        oracle = ref.OracleRiverine(mesh, {f"c{k}": inputs[k] for k in range(K)})
        n = plan.n_real
        rows = [np.empty(plan.n_face) for _ in range(K)]
        fl = [[np.empty(plan.n_edge) for _ in range(K)] for _ in range(3)]
        worst = 0.0
        for t in range(2):
            info = be.step(t)
            assert info.status == 0, (name, info.status)
            oracle.update()
            be.fetch_async(t + 1, rows, fl[0], fl[1], fl[2])
            be.fetch_wait()
            for k in range(K):
                want = oracle.constituent_dict[f"c{k}"].concentration[t + 1][:n]
                worst = max(worst, float(np.abs(rows[k][:n] - want).max() / np.abs(want).max()))
        be.mass_totals(0, 0, 2); be.flux_sums(0); be.volume_sums()
        be.close()
        assert worst < 1e-9, (name, worst)
        print(f"{name}: ok, max scaled |gpu - oracle| = {worst:.2e}", flush=True)
    print("SANITIZE_CASES_OK")


if __name__ == "__main__":
    main()
