"""CPU experiment (scipy emulation of the Gauss-Seidel sweep ORDER, no GPU): error factor per sweep for the library's
flow-aligned colourings (cwr_order_cells) with different colour counts, hints and weak-edge thresholds.  Results in
profiles/r01_notes.md (0.31 per sweep at 11 colours, 0.25 with the 0.1 threshold, ~0.1 floor from diffusion)."""
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla, time, sys
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[2]))
from clearwater_riverine_b200 import synthetic
from clearwater_riverine_b200.backend import order_cells
from oracle import reference_step as ref
side=300
T=30
plan = synthetic.make_plan(side, side, T, dt=30.0, tri_fraction=0.1, dry_fraction=0.02, courant=1.5, n_exact=100_000, seed=2)
D=0.1
adv,_,_,cdiff,dt = ref.derive_coefficients(plan.face_flow, plan.edge_velocity, plan.face_x, plan.face_y, plan.f1, plan.f2, D, plan.time_seconds)
mesh = ref.HydroMesh(plan.f1, plan.f2, plan.n_face, adv, cdiff, plan.edge_velocity, plan.volume, dt, D)
n=plan.n_real
hint = plan.face_flow[::1].mean(0)
def matrix(t):
    lhs = ref.LHS(mesh); lhs.update_values(mesh, t)
    A = lhs.to_csr(); A.sum_duplicates()
    d = A.diagonal()
    return sp.diags(1.0/d) @ A      # row scaled
def gs_rate(A, perm, sweeps=6):
    # GS sweeps from zero on A z = u in order perm; measure error reduction per sweep vs exact
    P = sp.csr_matrix((np.ones(n), (np.arange(n), perm)), shape=(n,n))   # new i <- old perm[i]
    Ap = (P @ A @ P.T).tocsr()
    Lo = sp.tril(Ap, 0).tocsr(); Up = sp.triu(Ap, 1).tocsr()
    rng = np.random.default_rng(0)
    u = rng.random(n)
    exact = spla.spsolve(Ap.tocsc(), u)
    z = np.zeros(n); errs=[]
    for s in range(sweeps):
        z = spla.spsolve_triangular(Lo, u - Up @ z, lower=True)
        errs.append(np.linalg.norm(z-exact)/np.linalg.norm(exact))
    return errs
def jac_rate(A, sweeps=12):
    N = sp.identity(n) - A
    rng = np.random.default_rng(0); u = rng.random(n)
    exact = spla.spsolve(A.tocsc(), u); z = np.zeros(n); errs=[]
    for s in range(sweeps):
        z = u + N @ z; errs.append(np.linalg.norm(z-exact)/np.linalg.norm(exact))
    return errs
for t in ():
    A = matrix(t)
    print('t',t,'jacobi per-sweep factor', (jac_rate(A)[-1]/jac_rate(A)[5])**(1/6))
    for nc in (11, 32, 64):
        p, cptr, nl = order_cells(plan.f1, plan.f2, plan.n_face, True, nc, hint)
        perm = np.argsort(p)    # old id of new row
        e = gs_rate(A, perm)
        print('   colours', nc, 'err after sweeps', ['%.1e'%x for x in e], 'factor/sweep', (e[-1]/e[1])**(1/4))
    # ideal: topological order of the ACTUAL flow at time t (hint = flow at t), many colours
    p, cptr, nl = order_cells(plan.f1, plan.f2, plan.n_face, True, 64, plan.face_flow[t])
    e = gs_rate(A, np.argsort(p)); print('   hint=flow[t], 64 colours', ['%.1e'%x for x in e])
    # steady part only as hint
print('---- thresholded hints')
absq = np.abs(hint)
cellmax = np.zeros(plan.n_face); np.maximum.at(cellmax, plan.f1, absq); np.maximum.at(cellmax, plan.f2, absq)
edge_ref = np.maximum(cellmax[plan.f1], np.where(plan.f2 < n, cellmax[plan.f2], 0))
for tau in (0.0, 0.1, 0.25, 0.4, 0.6):
    h2 = np.where(absq >= tau*edge_ref, hint, 0).astype(np.float32)
    for nc in (11, 16):
        p, cptr, nl = order_cells(plan.f1, plan.f2, plan.n_face, True, nc, h2)
        res=[]
        for t in (3, 10, 20):
            A = matrix(t); e = gs_rate(A, np.argsort(p)); res.append((e[-1]/e[1])**(1/4))
        print('tau', tau, 'colours', nc, 'levels', nl, 'factor/sweep at t=3,10,20:', ['%.3f'%x for x in res], 'first-sweep err', '%.2f'%e[0])
