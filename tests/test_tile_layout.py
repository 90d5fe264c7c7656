"""EXPERIMENTAL (precond_sweep = 2): host-side tiling of the tile-local Gauss-Seidel sweeps
(csrc/cwr_topology.cpp through cwr_tile_layout) -- structure, and a numpy emulation of exactly what k_precond_tile
does with it, used as a BiCGSTAB preconditioner on the oracle's matrix.  No device needed."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from clearwater_riverine_b200 import synthetic
from clearwater_riverine_b200.backend import tile_layout
from oracle import reference_step as ref

MASK31 = 0x3FFFFFFF
LATER, OUTSIDE, IDX = 0x8000, 0x4000, 0x1FFF


@pytest.fixture(scope="module")
def case():
    plan = synthetic.make_plan(90, 70, 5, tri_fraction=0.15, dry_fraction=0.02, seed=12)
    D = 0.1
    adv, _, _, cdiff, dt = ref.derive_coefficients(plan.face_flow, plan.edge_velocity, plan.face_x, plan.face_y,
                                                   plan.f1, plan.f2, D, plan.time_seconds)
    mesh = ref.HydroMesh(plan.f1, plan.f2, plan.n_face, adv, cdiff, plan.edge_velocity, plan.volume, dt, D)
    L = tile_layout(plan.f1, plan.f2, plan.n_face, 11, plan.face_flow.mean(0), 400, 1100, 4)
    return plan, mesh, L


def test_tiles_partition_the_rows_and_local_indices_are_consistent(case):
    plan, mesh, L = case
    n, nt, nc = plan.n_real, L["n_tiles"], L["n_colors"]
    tp, ep, rows, lcp, tell, ecol = L["tile_ptr"], L["ext_ptr"], L["ext_rows"], L["lcolor_ptr"], L["tile_ell"], L["ell_col"]
    assert tp[0] == 0 and tp[-1] == n and np.all(np.diff(tp) > 0) and np.diff(tp).max() <= 400
    assert ep[0] == 0 and ep[-1] == len(rows) and np.diff(ep).max() <= 1100
    g = rows & MASK31
    halo = rows < 0
    assert np.array_equal(np.sort(g[~halo]), np.arange(n))            # every row is core in exactly one tile
    colour_of = np.full(n, -1)
    for t in range(nt):
        base, m = ep[t], ep[t + 1] - ep[t]
        gl, hl = g[base:base + m], halo[base:base + m]
        core = gl[~hl]
        assert core.min() == tp[t] and core.max() == tp[t + 1] - 1 and len(core) == tp[t + 1] - tp[t]
        assert len(np.unique(gl)) == m
        assert lcp[t, 0] == 0 and lcp[t, -1] == m and np.all(np.diff(lcp[t]) >= 0)
        col_local = np.searchsorted(lcp[t], np.arange(m), side="right") - 1
        known = colour_of[gl] >= 0
        assert np.all(colour_of[gl][known] == col_local[known])       # a row has the same colour in every tile
        colour_of[gl] = col_local
        local_of = {int(r): l for l, r in enumerate(gl)}
        code = tell[base:base + m]
        nb = ecol[gl] & MASK31
        for l in range(0, m, 7):                                      # sample rows
            for w in range(L["W"]):
                j = int(nb[l, w])
                if code[l, w] & OUTSIDE:
                    assert j not in local_of
                else:
                    assert local_of[j] == (code[l, w] & IDX)
                assert bool(code[l, w] & LATER) == bool(ecol[gl[l], w] < 0)
    # colours separate coupled rows (padding entries point at the row itself)
    nb = ecol & MASK31
    real = nb != np.arange(n)[:, None]
    assert np.all((colour_of[nb] != colour_of[:, None])[real])
    # "later" = neighbour's colour >= the row's
    assert np.array_equal(ecol < 0, colour_of[nb] >= colour_of[:, None])


def emulate_tile_preconditioner(L, val, u, sweeps):
    """What k_precond_tile computes: per tile, `sweeps` multicolour Gauss-Seidel sweeps on core + halo rows from z = 0
    (rows outside the tile count as 0), core rows written back."""
    out = np.zeros_like(u)
    tp, ep, rows, lcp, tell = L["tile_ptr"], L["ext_ptr"], L["ext_rows"], L["lcolor_ptr"], L["tile_ell"]
    g, halo = rows & MASK31, rows < 0
    for t in range(L["n_tiles"]):
        base, m = ep[t], ep[t + 1] - ep[t]
        gl = g[base:base + m]
        ul, vl, code = u[gl], val[gl], tell[base:base + m]
        idx, outside, later = (code & IDX).astype(np.int64), (code & OUTSIDE) != 0, (code & LATER) != 0
        idx[outside] = 0
        z = np.zeros(m)
        for s in range(sweeps):
            for c in range(L["n_colors"]):
                lo, hi = lcp[t, c], lcp[t, c + 1]
                if lo == hi:
                    continue
                zz = z[idx[lo:hi]]
                zz[outside[lo:hi]] = 0.0
                if s == 0:
                    zz[later[lo:hi]] = 0.0
                z[lo:hi] = ul[lo:hi] - (vl[lo:hi] * zz).sum(axis=1)
        keep = ~halo[base:base + m]
        out[gl[keep]] = z[keep]
    return out


def test_emulated_tile_local_sweeps_precondition_bicgstab(case):
    plan, mesh, L = case
    n = plan.n_real
    lhs = ref.LHS(mesh); lhs.update_values(mesh, 2)
    A = lhs.to_csr(); A.sum_duplicates()
    A = (sp.diags(1.0 / A.diagonal()) @ A).tocsr()
    p = L["new_of_old"].astype(np.int64)
    P = sp.csr_matrix((np.ones(n), (p, np.arange(n))), shape=(n, n))          # new <- old
    An = (P @ A @ P.T).tocsr()
    nb = L["ell_col"] & MASK31
    val = np.asarray(An[np.arange(n)[:, None], nb].todense())
    val[nb == np.arange(n)[:, None]] = 0.0                                    # padding / the unit diagonal
    assert np.allclose(An @ np.ones(n), 1.0 + val.sum(axis=1))                # the ELL holds every off-diagonal
    rng = np.random.default_rng(1)
    xtrue = 50 + 50 * rng.random(n)
    b = An @ xtrue
    x0 = xtrue * (1 + 0.05 * rng.standard_normal(n))
    counts = {}
    for name, M in (("none", None),
                    ("tile-local, 5 sweeps", spla.LinearOperator((n, n), matvec=lambda u: emulate_tile_preconditioner(L, val, u, 5)))):
        it = [0]
        x, info = spla.bicgstab(An, b, x0=x0.copy(), rtol=1e-12, atol=0.0, M=M, maxiter=500,
                                callback=lambda xk: it.__setitem__(0, it[0] + 1))
        assert info == 0 and np.abs(x - xtrue).max() <= 1e-8 * np.abs(xtrue).max()
        counts[name] = it[0]
    assert counts["tile-local, 5 sweeps"] <= 4 and counts["none"] >= 8 * counts["tile-local, 5 sweeps"], counts
