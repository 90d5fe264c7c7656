"""Mesh file for topo_hash: int32 (n_real, n_face, n_edge), f1, f2, float32 time-mean face flow."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from clearwater_riverine_b200 import synthetic  # noqa: E402

out, side = sys.argv[1], int(sys.argv[2])
plan = synthetic.make_plan(side, side, 3, dt=30.0, tri_fraction=0.1, dry_fraction=0.02, courant=1.5, seed=2)
with open(out, "wb") as f:
    np.array([plan.n_real, plan.n_face, len(plan.f1)], np.int32).tofile(f)
    plan.f1.astype(np.int32).tofile(f)
    plan.f2.astype(np.int32).tofile(f)
    plan.face_flow.mean(0).astype(np.float32).tofile(f)
print(plan.n_real, "cells ->", out)
