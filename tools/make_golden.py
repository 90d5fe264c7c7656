#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Runs in the build container only (needs /root/reference).  The reference's own
ClearwaterRiverine class -- its HDF reader, WQVariableCalculator, Constituent
IC/BC ingestion, linalg.LHS / linalg.RHS, update() and _mass_flux() -- is
executed from /root/reference/src through oracle/refshim (xarray / h5py
stand-ins; see that package's docstring).  Each .npz holds the raw HEC-RAS
arrays, the derived coefficients and input arrays the reference computed, and
the concentrations / mass fluxes / per-step LHS + RHS it produced, so that

  * tests/test_oracle_golden.py can pin oracle/reference_step.py bit-for-bit, and
  * the GPU parity tests can run on the B200 box, where /root/reference is absent.

Usage:  python tools/make_golden.py            (rewrites tests/golden/)
"""
from __future__ import annotations

import contextlib
import io
import sys
import tempfile
import warnings
from pathlib import Path

import numpy as np
import pandas as pd
from scipy.sparse import csr_matrix

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle.refshim import load_reference  # noqa: E402

DATA = Path("/root/reference/tests/data/simple_test_cases")
OUT = ROOT / "tests" / "golden"

PLANS = {
    "p01": ("plan01_10x5", "p01"),
    "p02": ("plan02_2x1", "p02"),
    "p03": ("plan03_2x1", "p03"),
}


def _paths(plan):
    d, tag = PLANS[plan]
    base = DATA / d
    return (base / f"clearWaterTestCases.{tag}.hdf", base / f"cwr_initial_conditions_{tag}.csv",
            base / f"cwr_boundary_conditions_{tag}.csv")


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return fn(*a, **k)


def run_case(cwr, name, plan, diffusion, datetime_range, constituents, overrides=None, snapshot_steps=(0, 1, 5)):
    """constituents: {cname: (ic_csv, bc_csv)};  overrides: {step: {cname: array(n)}}"""
    hdf, _, _ = _paths(plan)
    cdict = {c: {"initial_conditions": str(ic), "boundary_conditions": str(bc), "units": "mg/L"}
             for c, (ic, bc) in constituents.items()}
    model = _quiet(cwr.ClearwaterRiverine, flow_field_file_path=str(hdf), diffusion_coefficient_input=diffusion,
                   constituent_dict=cdict, datetime_range=datetime_range)
    mesh = model.mesh
    T = len(mesh["time"])
    n = int(mesh.nreal) + 1
    names = list(constituents)
    snaps = {}
    for t in range(T - 1):
        upd = None
        if overrides and t in overrides:
            upd = {c: cwr.xr.DataArray(v) for c, v in overrides[t].items()}
        _quiet(model.update, upd)
        if t in snapshot_steps:
            A = csr_matrix((model.lhs.coef, (model.lhs.rows, model.lhs.cols)), shape=(n, n))
            A.sum_duplicates()
            A.sort_indices()
            snaps[f"A_indptr_{t}"] = A.indptr.astype(np.int64)
            snaps[f"A_indices_{t}"] = A.indices.astype(np.int64)
            snaps[f"A_data_{t}"] = A.data.copy()
            for c in names:
                snaps[f"b_{c}_{t}"] = model.constituent_dict[c].b.vals.copy()
    assert model.time_step == T - 1
    tsec = (mesh["time"].values - mesh["time"].values[0]) / np.timedelta64(1, "s")
    bd = model.boundary_data
    out = dict(
        # raw HEC-RAS arrays
        f1=np.asarray(mesh["edges_face1"].values, dtype=np.int32),
        f2=np.asarray(mesh["edges_face2"].values, dtype=np.int32),
        face_x=np.asarray(mesh["face_x"].values, dtype=np.float64),
        face_y=np.asarray(mesh["face_y"].values, dtype=np.float64),
        time_seconds=np.asarray(tsec, dtype=np.float64),
        face_flow=np.asarray(mesh["face_flow"].values),
        edge_velocity=np.asarray(mesh["edge_velocity"].values),
        volume=np.asarray(mesh["volume"].values),
        bc_names=np.array([str(x) for x in bd["Name"]]),
        bc_faces=np.asarray(bd["Face Index"], dtype=np.int64),
        diffusion_coefficient=np.float64(diffusion),
        # derived by the reference (utilities.py:513-541)
        adv=np.asarray(mesh["advection_coeff"].values),
        area=np.asarray(mesh["edge_vertical_area"].values),
        dist=np.asarray(mesh["face_to_face_dist"].values),
        cdiff=np.asarray(mesh["coeff_to_diffusion"].values),
        dt=np.asarray(mesh["dt"].values),
        constituents=np.array(names),
        snapshot_steps=np.array(sorted(s for s in snapshot_steps if s < T - 1)),
        **snaps,
    )
    assert out["adv"].dtype == np.float32 and out["cdiff"].dtype == np.float64, (out["adv"].dtype, out["cdiff"].dtype)
    for c in names:
        con = model.constituent_dict[c]
        out[f"input_{c}"] = con.input_array.copy()
        out[f"conc_{c}"] = np.asarray(mesh[c].values).copy()
        out[f"advflux_{c}"] = con.advection_mass_flux.copy()
        out[f"diffflux_{c}"] = con.diffusion_mass_flux.copy()
        out[f"totflux_{c}"] = con.total_mass_flux.copy()
    if overrides:
        steps = sorted(overrides)
        out["override_steps"] = np.array(steps)
        for t in steps:
            for c, v in overrides[t].items():
                out[f"override_{c}_{t}"] = np.asarray(v, dtype=np.float64)
    OUT.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / f"{name}.npz", **out)
    sz = (OUT / f"{name}.npz").stat().st_size
    print(f"{name}: T={T} n={n} F={len(mesh['nface'])} E={len(mesh['nedge'])}  "
          f"c_end[{names[0]}] range [{np.nanmin(out['conc_' + names[0]][-1][:n]):.12g}, "
          f"{np.nanmax(out['conc_' + names[0]][-1][:n]):.12g}]  {sz / 1024:.0f} KiB")


def main():
    cwr = load_reference()
    cwr.xr = sys.modules["xarray"]
    rng = np.random.default_rng(20230101)
    tmp = Path(tempfile.mkdtemp(prefix="golden_"))

    # --- the reference's own test cases: IC == BC == 100, D = 0.01 (tests/test_final_mass.py:22-27) ---
    _, ic2, bc2 = _paths("p02")
    run_case(cwr, "p02_uniform100", "p02", 0.01, None, {"tracer": (ic2, bc2)}, snapshot_steps=(0, 1, 5, 23))
    _, ic1, bc1 = _paths("p01")
    run_case(cwr, "p01_uniform100", "p01", 0.01, (0, 300), {"tracer": (ic1, bc1)}, snapshot_steps=(0, 1, 50, 299))
    _, ic3, bc3 = _paths("p03")
    run_case(cwr, "p03_uniform100", "p03", 0.01, (0, 400), {"tracer": (ic3, bc3)}, snapshot_steps=(0, 1, 200))

    # --- non-trivial parity case: random IC, distinct BC levels, two constituents, overrides ---
    ic_rand = tmp / "ic_rand.csv"
    pd.DataFrame({"Cell_Index": np.arange(50), "Concentration": 100.0 * (1.0 + rng.random(50))}).to_csv(ic_rand, index=False)
    bc = pd.read_csv(bc1)
    bc_rand = tmp / "bc_rand.csv"
    bc2_df = bc.copy()
    up = bc2_df["RAS2D_TS_Name"] == "US_Flow"
    bc2_df.loc[up, "Concentration"] = 250.0 + 50.0 * np.sin(np.arange(up.sum()) / 3.0)
    bc2_df.loc[~up, "Concentration"] = 40.0
    bc2_df.to_csv(bc_rand, index=False)
    overrides = {
        7: {"second": 100.0 * (1.0 + rng.random(50))},
        120: {"first": 50.0 + rng.random(50), "second": 10.0 * rng.random(50)},
    }
    run_case(cwr, "p01_random_two", "p01", 0.05, (2000, 2300), {"first": (ic_rand, bc_rand), "second": (ic1, bc1)},
             overrides=overrides, snapshot_steps=(0, 7, 120, 299))

    # --- D == 0 switches the RHS ghost diffusion off (linalg.py:390) ---
    run_case(cwr, "p01_nodiffusion", "p01", 0.0, (5000, 5100), {"first": (ic_rand, bc_rand)}, snapshot_steps=(0, 50))


if __name__ == "__main__":
    main()
