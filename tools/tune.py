#!/usr/bin/env python
"""Solver-option sweep on the bench workload (GPU only): ms/step, iterations and per-kernel-family times.

  python tools/tune.py [--workload 1m16] [--steps 6] [--scale 1.0] "precond_steps=8" "precond_steps=16,precond_sweep=1" ...

Each positional argument is one comma-separated cwr_options override set.  Development tool: prints one
line per set; bench.py is the measurement of record.
"""
from __future__ import annotations

import argparse
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="1m16")
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("sets", nargs="*", default=["precond_steps=8"])
    args = ap.parse_args()
    import bench
    from clearwater_riverine_b200 import TransportBackend, synthetic
    T = args.warmup + 2 * args.steps + 1
    plan, K = bench.workload_plan(args.workload, T, seed=2, scale=args.scale)
    inputs = synthetic.make_inputs(plan, K, seed=2)
    dt = np.append(np.diff(plan.time_seconds), np.nan)
    ref_state = None
    for spec in args.sets:
        opts = {}
        for kv in filter(None, spec.split(",")):
            key, val = kv.split("=", 1)
            if key.startswith("CWR_"):          # development switches read from the environment at create
                import os
                os.environ[key] = val
                continue
            opts[key] = float(val) if key == "rtol" else int(val)
        t_c = time.perf_counter()
        be = TransportBackend(plan.f1, plan.f2, plan.n_face, T, K, bench.DIFFUSION, device=0, mass_flux=0, **opts)
        t_c = time.perf_counter() - t_c
        be.set_geometry(plan.face_x, plan.face_y)
        be.set_hydro_raw(0, plan.face_flow, plan.edge_velocity, plan.volume, dt)
        for k in range(K):
            be.set_inputs(k, inputs[k])
        for t in range(args.warmup):
            be.step(t)
        _, i0 = be.counters()
        sw0 = be.solver_stats()[0]
        t0 = time.perf_counter()
        info = be.run(args.warmup, args.warmup + args.steps)
        wall = (time.perf_counter() - t0) / args.steps * 1e3
        _, i1 = be.counters()
        sw1, fb, nstrips, maxnbr = be.solver_stats()
        be.profile(1)
        for t in range(args.warmup + args.steps, args.warmup + 2 * args.steps):
            be.step(t)
        prof = be.profile(0)
        fam = {k: round(v[0] / args.steps, 3) for k, v in prof.items() if v[1]}
        per = {k: round(v[0] / v[1] * 1e3, 1) for k, v in prof.items() if v[1]}
        state = be.get_state_all(args.warmup + args.steps)
        if ref_state is None:
            ref_state = state
        dev = float(np.nanmax(np.abs(state - ref_state)) / np.nanmax(np.abs(ref_state)))
        o = be.options
        print(f"[{spec}] {wall:.3f} ms/step  iters/step {(i1 - i0) / args.steps:.2f} sweeps/step {(sw1 - sw0) / args.steps:.2f} fallbacks {fb} "
              f"solver {o.solver} sync {o.precond_sync} colours {o.precond_colors} strips {nstrips} (max nbrs {maxnbr})  status {info.status} relres {info.max_relres:.2e} "
              f"dev_vs_first {dev:.2e} create {t_c:.2f}s\n    ms/step by family: {fam}\n    us/launch: {per}", flush=True)
        be.close()


if __name__ == "__main__":
    main()
