"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle and the
golden vectors of the unmodified reference.  Tolerance (BASELINE.json north_star): concentrations
within rtol 1e-9 per step of the reference's spsolve path on the same inputs."""
import numpy as np
import pytest

from oracle import reference_step as ref
from tests.helpers import GOLDEN_CASES, golden_mesh, golden_overrides, load_golden

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def close(got, want, rtol=RTOL, what=""):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert np.array_equal(nan_g, nan_w), f"{what}: NaN pattern differs"
    fin = ~nan_w
    if not fin.any():
        return 0.0
    scale = np.abs(want[fin]).max()
    err = np.abs(got[fin] - want[fin]).max()
    assert err <= rtol * max(scale, 1e-300), f"{what}: max abs err {err:.3e} vs scale {scale:.3e}"
    return err / max(scale, 1e-300)


def flux_close(be, k, t, mesh, c_next, want_adv, want_diff, want_tot, what):
    """Fluxes are products of coefficients with concentrations (and, for diffusion, with a DIFFERENCE of
    concentrations, which cancels): the error bound is rtol * |coefficient| * |c| * dt."""
    fa, fd, ft = be.get_mass_flux(k, t)
    cmax = np.nanmax(np.abs(c_next))
    scale = max(np.abs(mesh.adv[t]).max(), np.abs(mesh.cdiff[t]).max()) * cmax * mesh.dt[t]
    for got, want, nm in ((fa, want_adv, "advection"), (fd, want_diff, "diffusion"), (ft, want_tot, "total")):
        assert np.array_equal(np.isnan(got), np.isnan(want)), f"{what}: {nm} flux NaN pattern"
        fin = ~np.isnan(want)
        if fin.any():
            err = np.abs(got[fin] - want[fin]).max()
            assert err <= RTOL * scale, f"{what}: {nm} flux err {err:.3e} vs scale {scale:.3e}"


def make_backend(mesh, inputs, **opt):
    from clearwater_riverine_b200 import TransportBackend
    be = TransportBackend(mesh.f1, mesh.f2, mesh.n_face, mesh.n_time, len(inputs), mesh.diffusion_coefficient, **opt)
    be.set_hydro(0, mesh.adv, mesh.cdiff, mesh.vel, mesh.vol, mesh.dt)
    for k, a in enumerate(inputs):
        be.set_inputs(k, a)
    return be


@pytest.mark.parametrize("case", GOLDEN_CASES)
@pytest.mark.parametrize("reorder,solver_path", [(1, 1), (0, 1), (1, 2)])
def test_golden_free_running(case, reorder, solver_path):
    """Whole trajectory against what the unmodified reference produced (tests/golden), through both
    solver paths: 1 = multi-CTA kernels + host-driven iteration, 2 = one CTA per constituent."""
    g = load_golden(case)
    mesh = golden_mesh(g)
    names = [str(c) for c in g["constituents"]]
    be = make_backend(mesh, [g[f"input_{c}"] for c in names], reorder=reorder, solver_path=solver_path)
    overrides = golden_overrides(g)
    snaps = set(int(s) for s in g["snapshot_steps"])
    n = mesh.n
    from scipy.sparse import csr_matrix
    for t in range(mesh.n_time - 1):
        for c, v in overrides.get(t, {}).items():
            be.set_state(names.index(c), t, v)
        if t > 0:   # row t as the reference's Dataset holds it (an override rewrites history, transport.py:233-236)
            for k, c in enumerate(names):
                close(be.get_state(k, t), g[f"conc_{c}"][t], RTOL, f"{case} c[{t}] {c}")
        info = be.step(t)
        assert info.status == 0, (t, info.status, info.max_relres)
        if t in snaps:
            A = be.get_lhs()
            Ag = csr_matrix((g[f"A_data_{t}"], g[f"A_indices_{t}"], g[f"A_indptr_{t}"]), shape=(n, n))
            assert np.array_equal(A.indptr, Ag.indptr) and np.array_equal(A.indices, Ag.indices)
            close(A.data, Ag.data, 1e-14, f"LHS step {t}")
            for k, c in enumerate(names):
                close(be.get_rhs(k), g[f"b_{c}_{t}"], 1e-13 if t == 0 else RTOL, f"RHS {c} step {t}")
        for k, c in enumerate(names):
            flux_close(be, k, t, mesh, g[f"conc_{c}"][t + 1], g[f"advflux_{c}"][t], g[f"diffflux_{c}"][t],
                       g[f"totflux_{c}"][t], f"{case} step {t} {c}")
    for k, c in enumerate(names):
        close(be.get_state(k, mesh.n_time - 1), g[f"conc_{c}"][mesh.n_time - 1], RTOL, f"{case} last row {c}")
        close(be.get_state(k, 0), g[f"conc_{c}"][0], 1e-15, "row 0")     # the IC row: zeros (not NaN) where unset
    be.close()


@pytest.mark.parametrize("case", ["p01_random_two", "p02_uniform100"])
def test_golden_teacher_forced_per_step(case):
    """Per-step parity on the same inputs: start every step from the reference's own c[t]."""
    g = load_golden(case)
    mesh = golden_mesh(g)
    names = [str(c) for c in g["constituents"]]
    be = make_backend(mesh, [g[f"input_{c}"] for c in names])
    overrides = golden_overrides(g)
    n = mesh.n
    for t in range(mesh.n_time - 1):
        for k, c in enumerate(names):
            state = g[f"conc_{c}"][t][:n].copy()
            if t in overrides and c in overrides[t]:
                state = overrides[t][c]
            be.set_state(k, t, state)
        assert be.step(t).status == 0
        for k, c in enumerate(names):
            if (t + 1) in overrides and c in overrides[t + 1]:
                continue               # the reference overwrote this history row with the override
            close(be.get_state(k, t + 1)[:n], g[f"conc_{c}"][t + 1][:n], RTOL, f"{case} step {t} {c}")
    be.close()


def test_raw_inputs_derived_on_device_match_reference_coefficients():
    """N1: adv / cdiff computed by the device from Face Flow / Face Velocity / coordinates
    (utilities.py:513-541) give the same trajectory as the reference-derived arrays."""
    from clearwater_riverine_b200 import TransportBackend
    g = load_golden("p01_random_two")
    mesh = golden_mesh(g)
    names = [str(c) for c in g["constituents"]]
    be = TransportBackend(mesh.f1, mesh.f2, mesh.n_face, mesh.n_time, len(names), mesh.diffusion_coefficient)
    be.set_geometry(g["face_x"], g["face_y"])
    be.set_hydro_raw(0, g["face_flow"], g["edge_velocity"], g["volume"], g["dt"])
    for k, c in enumerate(names):
        be.set_inputs(k, g[f"input_{c}"])
    for t in range(6):                 # the golden run overrides a constituent at step 7
        assert be.step(t).status == 0
        if t == 0:
            from scipy.sparse import csr_matrix
            Ag = csr_matrix((g["A_data_0"], g["A_indices_0"], g["A_indptr_0"]), shape=(mesh.n, mesh.n))
            close(be.get_lhs().data, Ag.data, 1e-14, "LHS from raw inputs")
        for k, c in enumerate(names):
            close(be.get_state(k, t + 1), g[f"conc_{c}"][t + 1], RTOL, f"raw-input path c[{t + 1}]")
    be.close()


def synthetic_case(nx, ny, T, K, seed, D=0.1, **kw):
    from clearwater_riverine_b200 import synthetic
    plan = synthetic.make_plan(nx, ny, T, seed=seed, **kw)
    adv, _, _, cdiff, dt = ref.derive_coefficients(plan.face_flow, plan.edge_velocity, plan.face_x, plan.face_y,
                                                   plan.f1, plan.f2, D, plan.time_seconds)
    mesh = ref.HydroMesh(plan.f1, plan.f2, plan.n_face, adv, cdiff, plan.edge_velocity, plan.volume, dt, D)
    inputs = synthetic.make_inputs(plan, K, seed=seed)
    return plan, mesh, inputs


def run_against_oracle(mesh, inputs, steps, rtol=RTOL, **opt):
    K = len(inputs)
    be = make_backend(mesh, list(inputs), **opt)
    oracle = ref.OracleRiverine(mesh, {f"c{k}": inputs[k] for k in range(K)})
    worst = 0.0
    for t in range(steps):
        info = be.step(t)
        assert info.status == 0, (t, info.status, info.iterations, info.max_relres)
        oracle.update()
        for k in range(K):
            con = oracle.constituent_dict[f"c{k}"]
            worst = max(worst, close(be.get_state(k, t + 1), con.concentration[t + 1], rtol, f"step {t} k {k}"))
            flux_close(be, k, t, mesh, con.concentration[t + 1], con.advection_mass_flux[t], con.diffusion_mass_flux[t],
                       con.total_mass_flux[t], f"step {t} k {k}")
    be.close()
    return worst


@pytest.mark.parametrize("solver_path", [1, 2])
@pytest.mark.parametrize("K", [1, 2, 3, 5, 16, 33])
def test_synthetic_mesh_all_column_widths(K, solver_path):
    """Quad-dominant shuffled mesh with dry cells; every lanes-per-row instantiation of the kernels."""
    _, mesh, inputs = synthetic_case(40, 25, 8, K, seed=K, dry_fraction=0.02)
    run_against_oracle(mesh, inputs, 7, solver_path=solver_path)


@pytest.mark.parametrize("m", [1, 2, 5])
def test_preconditioner_depths(m):
    """m-step Jacobi polynomial preconditioner: m = 1 is plain diagonal (Jacobi) preconditioning."""
    _, mesh, inputs = synthetic_case(40, 25, 6, 4, seed=17, dry_fraction=0.02)
    for path in (1, 2):
        run_against_oracle(mesh, inputs, 5, solver_path=path, precond_steps=m)


@pytest.mark.parametrize("K", [1, 2, 3, 4, 16, 33, 128])
@pytest.mark.parametrize("opts", [dict(precond_sweep=0, precond_precision=64, precond_steps=4),
                                  dict(precond_sweep=0, precond_steps=8),
                                  dict(precond_sweep=1, precond_steps=3),
                                  dict(precond_sweep=1, precond_steps=2, precond_colors=7),
                                  dict(precond_sweep=1, precond_steps=4, precond_precision=64, precond_colors=32)])
def test_preconditioner_variants(K, opts):
    """fp32 / fp64 sweeps, Jacobi steps / flow-aligned multicolour Gauss-Seidel (persistent kernel with grid
    barriers): the preconditioner changes the iteration count, never the converged answer."""
    if K == 128 and opts.get("precond_precision") == 64 and opts["precond_sweep"] == 0:
        pytest.skip("covered by K = 33")
    _, mesh, inputs = synthetic_case(40, 25, 6, K, seed=100 + K, dry_fraction=0.02)
    run_against_oracle(mesh, inputs, 5, solver_path=1, **opts)


SOLVER_VARIANTS = [dict(solver=1, precond_sync=1), dict(solver=1, precond_sync=2), dict(solver=2, precond_sync=1),
                   dict(solver=2, precond_sync=2), dict(solver=2, precond_sync=2, precond_precision=64),
                   dict(solver=2, precond_sync=3), dict(solver=1, precond_sync=3), dict(solver=2, precond_sync=3, precond_precision=64),
                   dict(solver=2, precond_sync=3, precond_colors=5),
                   dict(solver=2, precond_sync=2, precond_colors=5), dict(solver=2, precond_sweep=0, precond_steps=6),
                   dict(solver=2, precond_sync=3, precond_colors=40),
                   dict(solver=2, precond_sync=4), dict(solver=1, precond_sync=4), dict(solver=2, precond_sync=4, precond_colors=5),
                   dict(solver=2, precond_sync=4, precond_colors=40)]


@pytest.mark.parametrize("K", [1, 3, 16])
@pytest.mark.parametrize("opts", SOLVER_VARIANTS, ids=lambda o: "-".join(f"{k}{v}" for k, v in o.items()))
def test_solver_and_sweep_kernel_variants(K, opts):
    """Large-mesh path on a 12k-cell mesh (23 strips): BiCGSTAB / defect correction with the sweeps as the solver,
    grid-barrier / neighbour-synchronised / software-pipelined Gauss-Seidel kernel, Jacobi steps --
    the solver and the sweep kernel change the work, never the converged answer."""
    _, mesh, inputs = synthetic_case(120, 90, 6, K, seed=200 + K, dry_fraction=0.02)
    be = make_backend(mesh, list(inputs), solver_path=1, **opts)
    assert be.options.solver == opts["solver"]
    if opts.get("precond_sweep", 1) == 1:
        sync = opts["precond_sync"]
        if sync == 4 and not (K == 16 and opts.get("precond_precision", 32) == 32):
            sync = 3               # the TMA-ring kernel is written for fp32 sweeps over 16 constituents
        if sync >= 3 and (K * (4 if opts.get("precond_precision", 32) == 32 else 8)) % 16 != 0:
            sync = 2               # the pipelined kernel moves 16-byte packs
        assert be.options.precond_sync == sync
        assert be.solver_stats()[2] == (23 if sync >= 2 else 0)
    be.close()
    run_against_oracle(mesh, inputs, 5, solver_path=1, **opts)


def test_strip_kernel_with_every_cta_and_sync_variants_agree():
    """170k cells: one strip per resident CTA (2 x 148).  The grid-barrier and the neighbour-synchronised kernels do
    the same arithmetic on differently ordered rows: answers agree far inside rtol, the sweep counts are equal."""
    _, mesh, inputs = synthetic_case(410, 380, 4, 4, seed=77, dry_fraction=0.02)
    outs, sweeps = [], []
    for sync in (1, 2, 3):
        be = make_backend(mesh, list(inputs), solver_path=1, solver=2, precond_sync=sync)
        for t in range(3):
            info = be.step(t)
            assert info.status == 0 and info.max_relres <= 1e-13 and info.sweeps > 0
        outs.append(be.get_state_all(3))
        sweeps.append(be.solver_stats()[0])
        if sync >= 2:
            assert be.solver_stats()[2] >= 148 and be.solver_stats()[1] == 0
        be.close()
    close(outs[1], outs[0], 1e-11, "neighbour-synchronised vs grid-barrier sweeps")
    assert np.array_equal(outs[2], outs[1]), "the pipelined strip kernel does the same arithmetic on the same rows"
    assert sweeps[0] == sweeps[1] == sweeps[2], sweeps
    oracle = ref.OracleRiverine(mesh, {f"c{k}": inputs[k] for k in range(2)})      # (two of the four columns: SuperLU takes seconds each)
    for _ in range(3):
        oracle.update()
    for k in range(2):
        close(outs[1][k], oracle.constituent_dict[f"c{k}"].concentration[3][:mesh.n], RTOL, f"170k cells k{k}")


def test_tma_ring_strip_kernel_is_bitwise_the_pipelined_one():
    """170k cells x 16 constituents, one strip per resident CTA: k_gs_tma (precond_sync = 4, the default for fp32 sweeps
    over 16 constituents on one rank: operand streams by cp.async.bulk into a per-warp ring, gathers in registers) runs
    k_gs_strip's schedule and arithmetic -- same bits, same sweep counts; and the colour balancing (no strip colour above
    one pass of the CTA) leaves no overflow rows."""
    _, mesh, inputs = synthetic_case(410, 380, 4, 16, seed=78, dry_fraction=0.02)
    outs, sweeps = [], []
    for sync in (3, 4, 0):
        be = make_backend(mesh, list(inputs), solver_path=1, solver=2, precond_sync=sync)
        assert be.options.precond_sync == (sync or 4)
        for t in range(3):
            info = be.step(t)
            assert info.status == 0 and info.max_relres <= 1e-13 and info.sweeps > 0
        outs.append(be.get_state_all(3))
        sweeps.append(be.solver_stats()[0])
        assert be.solver_stats()[2] >= 148 and be.solver_stats()[1] == 0
        be.close()
    assert np.array_equal(outs[1], outs[0]) and np.array_equal(outs[2], outs[0])
    assert sweeps[0] == sweeps[1] == sweeps[2], sweeps
    oracle = ref.OracleRiverine(mesh, {"c0": inputs[0]})
    for _ in range(3):
        oracle.update()
    close(outs[1][0], oracle.constituent_dict["c0"].concentration[3][:mesh.n], RTOL, "170k cells x 16, TMA-ring kernel, k0")


def test_defect_correction_falls_back_to_bicgstab_when_the_sweeps_diverge():
    """Velocity / flow sign disagreement on 5 % of the faces makes `area = flow / velocity` negative there: with D = 2
    the matrix has negative diagonal entries and Gauss-Seidel sweeps diverge (spectral radius ~3.4) while spsolve --
    and BiCGSTAB -- still solve it.  The defect-correction solver must notice and hand over."""
    from clearwater_riverine_b200 import synthetic
    plan = synthetic.make_plan(60, 50, 6, seed=41, dry_fraction=0.02)
    vel = plan.edge_velocity.copy()
    vel[:, np.random.default_rng(1).random(plan.n_edge) < 0.05] *= -1
    D = 2.0
    adv, _, _, cdiff, dt = ref.derive_coefficients(plan.face_flow, vel, plan.face_x, plan.face_y, plan.f1, plan.f2, D, plan.time_seconds)
    mesh = ref.HydroMesh(plan.f1, plan.f2, plan.n_face, adv, cdiff, vel, plan.volume, dt, D)
    inputs = synthetic.make_inputs(plan, 2, seed=41)
    be = make_backend(mesh, list(inputs), solver_path=1, solver=2)
    oracle = ref.OracleRiverine(mesh, {f"c{k}": inputs[k] for k in range(2)})
    for t in range(3):
        info = be.step(t)
        assert info.status == 0, (t, info.status, info.iterations, info.max_relres)
        oracle.update()
        for k in range(2):
            close(be.get_state(k, t + 1), oracle.constituent_dict[f"c{k}"].concentration[t + 1], 1e-8, f"fallback k{k} t{t}")
    assert be.solver_stats()[1] >= 1, "the sweeps diverge on this matrix: the solver should have fallen back"
    be.close()


def test_nan_boundary_and_zero_rhs_on_the_large_path():
    """NaN boundary value -> the column is NaN like spsolve's result (CWR_ENAN); a constituent with nothing set -> 0."""
    plan, mesh, inputs = synthetic_case(60, 40, 6, 3, seed=21)
    n = mesh.n
    ghost = np.nonzero(mesh.f2 >= n)[0]
    inputs[1][:] = 0.0
    inputs[2][3:, mesh.f2[ghost[1]]] = np.nan
    for solver in (1, 2):
        be = make_backend(mesh, list(inputs), solver_path=1, solver=solver)
        oracle = ref.OracleRiverine(mesh, {f"c{k}": inputs[k] for k in range(3)})
        for t in range(5):
            info = be.step(t)
            oracle.update()
            for k in range(2):
                close(be.get_state(k, t + 1), oracle.constituent_dict[f"c{k}"].concentration[t + 1], RTOL, f"k{k} t{t}")
            want = oracle.constituent_dict["c2"].concentration[t + 1]
            got = be.get_state(2, t + 1)
            if np.isnan(want[:n]).any():
                assert info.status == -5 and np.isnan(got[:n]).all()
            else:
                close(got, want, RTOL, f"k2 t{t}")
        be.close()


@pytest.mark.parametrize("opts", [dict(precond_sweep=0), dict(precond_sweep=0, precond_steps=3), dict(precond_steps=2),
                                  dict(precond_steps=6, precond_colors=16), dict(precond_colors=5)])
@pytest.mark.parametrize("shape", [(40, 25), (70, 50)])
def test_small_mesh_paths(shape, opts):
    """One CTA per constituent: k_solve_tiny (on chip, Gauss-Seidel; 1 and 4 rows per thread) and k_solve_small
    (Jacobi steps through L2)."""
    _, mesh, inputs = synthetic_case(shape[0], shape[1], 6, 3, seed=31, dry_fraction=0.02)
    run_against_oracle(mesh, inputs, 5, solver_path=2, **opts)


def test_on_chip_kernels_against_the_oracle_and_each_other(monkeypatch):
    """k_solve_chip (a thread keeps one row of every colour: solver_path 4) and k_solve_tiny (rows dealt by thread id:
    3) solve the same systems: both within rtol of the oracle, each bitwise repeatable; the wider mesh (colours of more than
    256 rows) can only take k_solve_tiny; NaN boundary values and an all-zero right-hand side behave alike."""
    _, mesh, inputs = synthetic_case(40, 25, 6, 3, seed=33, dry_fraction=0.02)
    inputs = [a.copy() for a in inputs]
    inputs[1][:] = 0.0                                   # zero initial state, zero boundary: b == 0
    states = {}
    for env, want_path in ((None, 4), ("1", 3)):
        if env is None:
            monkeypatch.delenv("CWR_TINY_KERNEL", raising=False)
        else:
            monkeypatch.setenv("CWR_TINY_KERNEL", env)
        for rep in range(2):
            be = make_backend(mesh, list(inputs), solver_path=2)
            assert be.options.solver_path == want_path
            oracle = ref.OracleRiverine(mesh, {f"c{k}": inputs[k] for k in range(3)})
            for t in range(5):
                assert be.step(t).status == 0
                oracle.update()
                for k in range(3):
                    close(be.get_state(k, t + 1), oracle.constituent_dict[f"c{k}"].concentration[t + 1], RTOL, f"path {want_path} t{t} k{k}")
            final = be.get_state_all(5)
            be.close()
            if rep == 0:
                states[want_path] = final
            else:
                assert np.array_equal(final, states[want_path], equal_nan=True), "not bitwise repeatable"
    close(states[4], states[3], 1e-11, "chip vs tiny")
    monkeypatch.delenv("CWR_TINY_KERNEL", raising=False)
    _, wide, winputs = synthetic_case(70, 50, 4, 2, seed=34)
    be = make_backend(wide, list(winputs), solver_path=2)
    assert be.options.solver_path == 3
    be.close()
    for colours, want_path in ((5, 4), (8, 4), (13, 4), (16, 3)):          # 8 / 12 / 14 colour slots per thread; too many colours
        be = make_backend(mesh, list(inputs), solver_path=2, precond_colors=colours)
        assert be.options.solver_path == want_path, (colours, be.options.solver_path)
        be.close()
        run_against_oracle(mesh, inputs, 4, solver_path=2, precond_colors=colours)


@pytest.mark.parametrize("precision", [32, 64])
def test_on_chip_kernel_sweep_precisions(precision):
    """k_solve_chip with fp32 sweeps inside the fp64 BiCGSTAB (precond_precision = 32, the default) and with fp64 sweeps:
    within rtol of the oracle at 8 / 12 / 14 colour slots and with few sweeps, in units 1e-20 and 1e+20 (the fp32 sweeps
    scale what they are given); bitwise repeatable."""
    _, mesh, inputs = synthetic_case(40, 25, 6, 3, seed=35, dry_fraction=0.02)
    inputs = [a.copy() for a in inputs]
    inputs[1] *= 1e-20                                   # tiny units: the fp32 sweeps scale what they are given
    inputs[2] *= 1e+20
    finals = []
    for colours in (5, 12, 13, 12):
        be = make_backend(mesh, list(inputs), solver_path=2, precond_colors=colours, precond_precision=precision)
        assert be.options.solver_path == 4 and be.options.precond_precision == precision
        oracle = ref.OracleRiverine(mesh, {f"c{k}": inputs[k] for k in range(3)})
        for t in range(5):
            assert be.step(t).status == 0
            oracle.update()
            for k in range(3):
                close(be.get_state(k, t + 1), oracle.constituent_dict[f"c{k}"].concentration[t + 1], RTOL, f"fp{precision} colours {colours} t{t} k{k}")
        finals.append(be.get_state_all(5))
        be.close()
    assert np.array_equal(finals[1], finals[3], equal_nan=True), "not bitwise repeatable"
    run_against_oracle(mesh, inputs, 4, solver_path=2, precond_steps=3, precond_precision=precision)


def test_gauss_seidel_is_deterministic_and_hint_independent():
    """Same inputs, with and without the flow hint (different row orders): both within rtol of the oracle;
    two runs with the same order are bitwise identical."""
    plan, mesh, inputs = synthetic_case(36, 30, 6, 4, seed=23, dry_fraction=0.02)
    from clearwater_riverine_b200 import TransportBackend
    outs = []
    for hint in (True, True, False):
        be = TransportBackend(mesh.f1, mesh.f2, mesh.n_face, mesh.n_time, 4, mesh.diffusion_coefficient,
                              solver_path=1, precond_sweep=1, precond_steps=3)
        if not hint:
            be.set_flow_hint(np.zeros(len(mesh.f1), np.float32))
        be.set_hydro(0, mesh.adv, mesh.cdiff, mesh.vel, mesh.vol, mesh.dt)
        for k in range(4):
            be.set_inputs(k, inputs[k])
        for t in range(5):
            assert be.step(t).status == 0
        outs.append(np.stack([be.get_state(k, 5) for k in range(4)]))
        be.close()
    assert np.array_equal(outs[0], outs[1], equal_nan=True)
    close(outs[2], outs[0], RTOL, "hint vs no hint")


def test_run_many_steps_without_host_round_trips():
    """cwr_run on the small path queues every step's launches and synchronises once."""
    _, mesh, inputs = synthetic_case(30, 20, 12, 3, seed=19, dry_fraction=0.02)
    be = make_backend(mesh, list(inputs), solver_path=2)
    info = be.run(0, 11)
    assert info.status == 0 and 0 < info.max_relres <= 1e-13 and info.iterations > 0
    oracle = ref.OracleRiverine(mesh, {f"c{k}": inputs[k] for k in range(3)})
    for _ in range(11):
        oracle.update()
    for k in range(3):
        for t in (1, 6, 11):
            close(be.get_state(k, t), oracle.constituent_dict[f"c{k}"].concentration[t], RTOL, f"run k{k} t{t}")
    be.close()


def test_run_step_run_sequences_and_the_system_after_a_run():
    """cwr_run on the small path hands every step its parameters from one upload; cwr_step writes them per step.  Mixed
    sequences (run, step, run; a run of one step; a run over the rest) stay within rtol of the oracle at every time index,
    and the system read back after a run (cwr_get_lhs / cwr_get_rhs) is the last step's, equal to the oracle's."""
    _, mesh, inputs = synthetic_case(30, 20, 12, 2, seed=29, dry_fraction=0.02)
    be = make_backend(mesh, list(inputs), solver_path=2)
    oracle = ref.OracleRiverine(mesh, {f"c{k}": inputs[k] for k in range(2)})
    assert be.run(0, 4).status == 0
    assert be.step(4).status == 0
    assert be.run(5, 6).status == 0
    assert be.run(6, 11).status == 0
    n = mesh.n
    for t in range(11):
        oracle.update()
    for k in range(2):
        for t in range(1, 12):
            close(be.get_state(k, t), oracle.constituent_dict[f"c{k}"].concentration[t], RTOL, f"k{k} t{t}")
    lhs = ref.LHS(mesh); lhs.update_values(mesh, 10)
    A = lhs.to_csr(); A.sum_duplicates(); A.sort_indices()
    got = be.get_lhs(); got.sort_indices()
    assert np.array_equal(got.indptr, A.indptr) and np.array_equal(got.indices, A.indices)
    close(got.data, A.data, 1e-13, "LHS after the run")
    rhs = ref.RHS(mesh, inputs[1])
    rhs.update_values(oracle.constituent_dict["c1"].concentration[10][:n].copy(), mesh, 10)
    close(be.get_rhs(1), np.asarray(rhs.vals, dtype=np.float64), RTOL, "RHS after the run")
    be.close()


@pytest.mark.parametrize("opts", [dict(reorder=0), dict(keep_history=0), dict(solver_path=1, check_every=3),
                                  dict(hydro_capacity=2)])
def test_option_variants(opts):
    plan, mesh, inputs = synthetic_case(30, 30, 6, 4, seed=5, dry_fraction=0.01)
    if "hydro_capacity" in opts:
        from clearwater_riverine_b200 import TransportBackend
        be = TransportBackend(mesh.f1, mesh.f2, mesh.n_face, mesh.n_time, 4, mesh.diffusion_coefficient, **opts)
        for k in range(4):
            be.set_inputs(k, inputs[k])
        oracle = ref.OracleRiverine(mesh, {f"c{k}": inputs[k] for k in range(4)})
        be.set_hydro(0, mesh.adv[0:1], mesh.cdiff[0:1], mesh.vel[0:1], mesh.vol[0:1], mesh.dt[0:1])
        for t in range(5):     # stream one slice ahead, two resident
            be.set_hydro(t + 1, mesh.adv[t + 1:t + 2], mesh.cdiff[t + 1:t + 2], mesh.vel[t + 1:t + 2],
                         mesh.vol[t + 1:t + 2], mesh.dt[t + 1:t + 2])
            assert be.step(t).status == 0
            oracle.update()
            for k in range(4):
                close(be.get_state(k, t + 1), oracle.constituent_dict[f"c{k}"].concentration[t + 1], RTOL, "streamed")
        be.close()
    else:
        run_against_oracle(mesh, inputs, 5, **opts)


def test_adversarial_boundary_semantics():
    """Reference quirks: several inflowing ghost edges on one cell (last edge wins on the
    RHS while the LHS sums), zero BC = unset (ghost stays NaN), NaN BC propagates, velocity / flow sign
    disagreement on a ghost edge, D == 0."""
    plan, mesh, inputs = synthetic_case(12, 9, 6, 3, seed=21, tri_fraction=0.3)
    n = mesh.n
    ghost = np.nonzero(mesh.f2 >= n)[0]
    rng = np.random.default_rng(0)
    # make every ghost edge carry flow, with random direction per time step -> corner cells get 2 active edges
    mesh.adv[:, ghost] = (rng.random((mesh.n_time, len(ghost))) - 0.4).astype(np.float32) * 3
    mesh.vel[:, ghost] = mesh.adv[:, ghost] / 20.0
    flip = ghost[:: 5]
    mesh.vel[:, flip] *= -1                        # sign(vel) != sign(adv): LHS follows adv, RHS follows vel
    mesh.cdiff[:, ghost] = np.abs(mesh.adv[:, ghost]) * 0.01 + 0.003
    inputs[1][:, mesh.f2[ghost[::3]]] = 0.0        # unset BCs
    inputs[2][3:, mesh.f2[ghost[1]]] = np.nan      # a boundary that starts late (merge_asof NaN)
    be = make_backend(mesh, list(inputs))
    oracle = ref.OracleRiverine(mesh, {f"c{k}": inputs[k] for k in range(3)})
    for t in range(5):
        info = be.step(t)
        oracle.update()
        for k in range(2):
            close(be.get_state(k, t + 1), oracle.constituent_dict[f"c{k}"].concentration[t + 1], RTOL, f"adv k{k} t{t}")
        want = oracle.constituent_dict["c2"].concentration[t + 1]
        got = be.get_state(2, t + 1)
        if np.isnan(want[:n]).any():
            assert info.status == -5               # CWR_ENAN, and the column is NaN like spsolve's result
            assert np.isnan(got[:n]).all()
        else:
            close(got, want, RTOL, f"adv k2 t{t}")
    be.close()


def test_zero_diffusion_and_zero_rhs():
    plan, mesh, inputs = synthetic_case(16, 10, 5, 2, seed=3)
    mesh.diffusion_coefficient = 0.0
    mesh.cdiff[:] = 0.0
    inputs[1][:] = 0.0                              # nothing set at all: b == 0 -> c == 0
    run_against_oracle(mesh, inputs, 4)


def test_uniform_concentration_is_preserved_on_steady_flow():
    """The invariant behind the reference's tests/test_final_mass.py (IC == BC == 100)."""
    plan, mesh, _ = synthetic_case(50, 30, 12, 1, seed=9, unsteady=0.0, tidal=0.0, dry_fraction=0.0)
    n = mesh.n
    inp = np.zeros((mesh.n_time, mesh.n_face)); inp[0, :n] = 100.0; inp[:, n:] = 100.0
    be = make_backend(mesh, [inp])
    for t in range(11):
        assert be.step(t).status == 0
    c = be.get_state(0, 11)[:n]
    assert np.abs(c - 100.0).max() < 1e-3          # float32 continuity error of the hydrodynamics only
    m = be.mass_totals(0, 0, 11)
    assert abs(m.mass_end - 100.0 * m.vol_end) / m.mass_end < 1e-5
    be.close()


@pytest.mark.parametrize("opts", [dict(), dict(stream_hydro=True), dict(stream_hydro=True, solver_path=1, keep_history=0)])
def test_host_mirror_update_loop_matches_oracle(opts):
    """ClearwaterRiverine.update() with update_concentration, as a coupling loop drives it
    (examples/03_01_coupling_riverine_modules_nsm.ipynb cell 47); also with the hydrodynamics streamed slice by
    slice (slice t+2 prefetched on the upload stream while c[t+1] is copied back)."""
    from clearwater_riverine_b200 import ClearwaterRiverine
    plan, mesh, inputs = synthetic_case(20, 14, 7, 2, seed=4)
    model = ClearwaterRiverine.from_arrays(plan.f1, plan.f2, plan.face_x, plan.face_y, plan.time_seconds, plan.face_flow,
                                           plan.edge_velocity, plan.volume, 0.1, {"a": inputs[0], "b": inputs[1]}, **opts)
    oracle = ref.OracleRiverine(mesh, {"a": inputs[0], "b": inputs[1]})
    n = mesh.n
    rng = np.random.default_rng(1)
    for t in range(6):
        upd = {"b": 50.0 + rng.random(n)} if t in (2, 4) else None
        model.update(upd)
        oracle.update(upd)
        assert model.time_step == oracle.time_step
        for name in ("a", "b"):
            close(model.mesh[name][t + 1], oracle.constituent_dict[name].concentration[t + 1], RTOL, f"{name} {t}")
            close(model.mesh[name][t], oracle.constituent_dict[name].concentration[t], RTOL, f"{name} history {t}")
            close(model.constituent_dict[name].total_mass_flux[t], oracle.constituent_dict[name].total_mass_flux[t], 1e-7, "flux")
    model.finalize()


def test_bitwise_repeatable():
    """Deterministic reductions: two runs give identical bits."""
    _, mesh, inputs = synthetic_case(40, 40, 5, 4, seed=8)
    outs = []
    for _ in range(2):
        be = make_backend(mesh, list(inputs))
        for t in range(4):
            be.step(t)
        outs.append(be.get_state_all(4).copy())
        be.close()
    assert np.array_equal(outs[0], outs[1])


def test_large_mesh_properties():
    """Size-independent properties at a size the oracle cannot run quickly (250k cells x 4):
    A x = b residual through an independent product, and the column-sum conservation identity."""
    from clearwater_riverine_b200 import synthetic
    plan = synthetic.make_plan(500, 450, 4, seed=13, dry_fraction=0.02)
    D = 0.1
    adv, _, _, cdiff, dt = ref.derive_coefficients(plan.face_flow, plan.edge_velocity, plan.face_x, plan.face_y,
                                                   plan.f1, plan.f2, D, plan.time_seconds)
    mesh = ref.HydroMesh(plan.f1, plan.f2, plan.n_face, adv, cdiff, plan.edge_velocity, plan.volume, dt, D)
    inputs = synthetic.make_inputs(plan, 4, seed=13)
    be = make_backend(mesh, list(inputs))
    n = mesh.n
    info = be.step(0)
    assert info.status == 0
    A = be.get_lhs()
    # conservation: column j sums to vol[t+1,j]/dt (+1 if dry) + ghost-edge terms  
    colsum = np.asarray(A.sum(axis=0)).ravel()
    expect = mesh.vol[1][:n].astype(np.float64) / dt[0] + (mesh.vol[1][:n] == 0)
    gh = np.nonzero(mesh.f2 >= n)[0]
    np.add.at(expect, mesh.f1[gh], cdiff[0][gh] + np.maximum(adv[0][gh].astype(np.float64), 0.0))
    assert np.abs(colsum - expect).max() <= 1e-12 * np.abs(expect).max()
    for k in range(4):
        x = be.get_state(k, 1)[:n]
        b = be.get_rhs(k)
        r = A @ x - b
        assert np.linalg.norm(r) <= 1e-11 * np.linalg.norm(b)
    # the oracle's own assembly agrees on this mesh too
    lhs = ref.LHS(mesh); lhs.update_values(mesh, 0)
    Ao = lhs.to_csr(); Ao.sum_duplicates(); Ao.sort_indices()
    assert np.array_equal(Ao.indices, A.indices)
    assert np.abs(Ao.data - A.data).max() <= 1e-13 * np.abs(Ao.data).max()
    be.close()


def test_full_size_1m_x16_residual_linearity_and_repeatability():
    """BASELINE configs[2] at its full size (1M cells x 16 constituents), where the oracle's SuperLU would need
    ~50 s per constituent-step: size-independent properties instead.  (1) A x = b through an independent host
    product for every constituent; (2) linearity: the constituent whose inputs are the sum of two others is
    their sum; (3) the default preconditioner (fp32 flow-aligned Gauss-Seidel) and fp64 Jacobi steps give the
    same answer; (4) two runs are bitwise identical."""
    import bench
    from clearwater_riverine_b200 import TransportBackend, synthetic
    T = 3
    plan, K = bench.workload_plan("1m16", T, seed=2)
    assert plan.n_real == 1_000_000 and K == 16
    n = plan.n_real
    inputs = synthetic.make_inputs(plan, K, seed=2)
    inputs[2] = inputs[0] + inputs[1]
    dt = np.append(np.diff(plan.time_seconds), np.nan)
    hint = plan.face_flow.mean(axis=0)
    outs = {}
    for name, opts in (("gs", {}), ("gs2", {}), ("jacobi64", dict(precond_sweep=0, precond_precision=64, precond_steps=8))):
        be = TransportBackend(plan.f1, plan.f2, plan.n_face, T, K, bench.DIFFUSION, flow_hint=hint, **opts)
        be.set_geometry(plan.face_x, plan.face_y)
        be.set_hydro_raw(0, plan.face_flow, plan.edge_velocity, plan.volume, dt)
        for k in range(K):
            be.set_inputs(k, inputs[k])
        for t in range(2):
            info = be.step(t)
            assert info.status == 0 and info.max_relres <= 1e-13
        outs[name] = be.get_state_all(2)
        if name == "gs":
            A = be.get_lhs()                                # LHS of the last step, reference numbering
            for k in range(K):
                b = be.get_rhs(k)
                r = A @ outs[name][k] - b
                assert np.linalg.norm(r) <= 1e-11 * np.linalg.norm(b), k
        be.close()
    x = outs["gs"]
    scale = np.abs(x).max()
    assert np.abs(x[2] - (x[0] + x[1])).max() <= 1e-9 * scale                       # linearity
    assert np.abs(x - outs["jacobi64"]).max() <= 1e-9 * scale                       # preconditioner-independent
    assert np.array_equal(x, outs["gs2"])                                           # bitwise repeatable
