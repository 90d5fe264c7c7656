"""Shared test helpers: golden-vector loading and oracle construction."""
from pathlib import Path

import numpy as np

from oracle import reference_step as ref

GOLDEN = Path(__file__).resolve().parent / "golden"
GOLDEN_CASES = ["p02_uniform100", "p01_uniform100", "p03_uniform100", "p01_random_two", "p01_nodiffusion"]


def load_golden(name):
    return dict(np.load(GOLDEN / f"{name}.npz", allow_pickle=False))


def golden_mesh(g, derived=True):
    """HydroMesh from a golden file; derived=True uses the reference-computed adv/cdiff/dt."""
    if derived:
        adv, cdiff, dt = g["adv"], g["cdiff"], g["dt"]
    else:
        adv, _area, _dist, cdiff, dt = ref.derive_coefficients(
            g["face_flow"], g["edge_velocity"], g["face_x"], g["face_y"], g["f1"], g["f2"],
            float(g["diffusion_coefficient"]), g["time_seconds"])
    return ref.HydroMesh(f1=g["f1"], f2=g["f2"], n_face=g["volume"].shape[1], adv=adv, cdiff=cdiff,
                         vel=g["edge_velocity"], vol=g["volume"], dt=dt,
                         diffusion_coefficient=float(g["diffusion_coefficient"]))


def golden_overrides(g):
    out = {}
    if "override_steps" in g:
        for t in g["override_steps"]:
            out[int(t)] = {str(c): g[f"override_{c}_{t}"] for c in g["constituents"] if f"override_{c}_{t}" in g}
    return out


def same(a, b):
    """Bitwise equality with NaN == NaN."""
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)
