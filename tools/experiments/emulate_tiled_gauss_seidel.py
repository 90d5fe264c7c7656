"""CPU experiment for round 2 (scipy emulation, no GPU): a TILE-LOCAL multi-sweep Gauss-Seidel preconditioner --
restricted additive Schwarz with overlap: every tile of the mesh does S sweeps on its core cells + d halo layers from
z = 0 with nothing outside, only the core rows are kept -- compared with the global multicolour Gauss-Seidel the
library uses now.  Such a tile fits in one CTA's shared memory (24 x 24 cells + 4 halo layers = 1024 rows x 16
constituents fp32 = 163 KB), needs no grid barrier, and reads the matrix and u from HBM once per application.
Result on the 100k-cell benchmark-like mesh (printed below): 24 x 24 tiles + 4 halo layers, 5 sweeps: 3 BiCGSTAB
iterations against 2 for the global sweeps; + 8 layers and 10 sweeps: 2.  Recorded in DESIGN.md section 7."""
# Emulation of a tile-local multi-sweep Gauss-Seidel preconditioner (restricted additive Schwarz with overlap):
# every tile does S sweeps on its core + d halo layers from z = 0 with nothing outside; only core rows are kept.
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla, sys, time
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[2]))
from clearwater_riverine_b200 import synthetic
from clearwater_riverine_b200.backend import order_cells
from oracle import reference_step as ref
T=14
plan = synthetic.make_plan(300, 300, T, dt=30.0, tri_fraction=0.1, dry_fraction=0.02, courant=1.5, n_exact=100_000, seed=2)
D=0.1
adv,_,_,cdiff,dt = ref.derive_coefficients(plan.face_flow, plan.edge_velocity, plan.face_x, plan.face_y, plan.f1, plan.f2, D, plan.time_seconds)
mesh = ref.HydroMesh(plan.f1, plan.f2, plan.n_face, adv, cdiff, plan.edge_velocity, plan.volume, dt, D)
n=plan.n_real
hint = plan.face_flow.mean(0)
t=8
lhs = ref.LHS(mesh); lhs.update_values(mesh, t); A = lhs.to_csr(); A.sum_duplicates()
A = (sp.diags(1.0/A.diagonal()) @ A).tocsr()
inputs = synthetic.make_inputs(plan, 1, seed=2)
rhs = ref.RHS(mesh, inputs[0]); rhs.update_values(inputs[0][0][:n].copy(), mesh, t); b = rhs.vals / lhs.to_csr().diagonal() if False else None
rng = np.random.default_rng(0)
xtrue = 50 + 50*rng.random(n); b = A @ xtrue; x0 = xtrue * (1 + 0.05*rng.standard_normal(n))   # warm start ~5 % off
p, cptr, nl = order_cells(plan.f1, plan.f2, plan.n_face, True, 11, hint)   # p = new_of_old: global sweep position
internal = plan.f2 < n
adjA = sp.csr_matrix((np.ones(2*internal.sum()), (np.concatenate([plan.f1[internal], plan.f2[internal]]), np.concatenate([plan.f2[internal], plan.f1[internal]]))), shape=(n,n))

def make_global_gs(S):
    perm = np.argsort(p); P = sp.csr_matrix((np.ones(n), (np.arange(n), perm)), shape=(n,n))
    Ap = (P @ A @ P.T).tocsr(); Lo = sp.tril(Ap,0).tocsr(); Up = sp.triu(Ap,1).tocsr()
    def apply(u):
        up = P @ u; z = np.zeros(n)
        for s in range(S): z = spla.spsolve_triangular(Lo, up - Up @ z, lower=True)
        return P.T @ z
    return apply

def make_tiled_gs(S, tile, depth):
    cx, cy = plan.face_x[:n], plan.face_y[:n]
    dx = 10.0
    ti = (np.floor(cx/(dx*tile)).astype(int)), (np.floor(cy/(dx*tile)).astype(int))
    tid = ti[0]*10000 + ti[1]
    uniq, tile_of = np.unique(tid, return_inverse=True); nt = len(uniq)
    # membership matrix (tile x cell), grown by `depth` layers
    Mb = sp.csr_matrix((np.ones(n), (tile_of, np.arange(n))), shape=(nt, n))
    Mx = Mb.copy()
    for _ in range(depth):
        Mx = ((Mx + Mx @ adjA) > 0).astype(float).tocsr()
    Mx = Mx.tocoo()
    ext_tile, ext_cell = Mx.row, Mx.col                  # extended rows: (tile, cell)
    order = np.lexsort((p[ext_cell], ext_tile))          # within a tile: global sweep order
    ext_tile, ext_cell = ext_tile[order], ext_cell[order]
    ne = len(ext_cell)
    # local index of (tile, cell)
    key = ext_tile.astype(np.int64)*n + ext_cell
    sorter = np.argsort(key); key_sorted = key[sorter]
    Ac = A.tocoo()
    # for every ext row r=(tile,i): entries A[i,j] with (tile,j) in ext
    rows_of_cell = {}
    # vectorised: expand each ext row's nonzeros
    indptr, indices, data = A.indptr, A.indices, A.data
    cnt = indptr[ext_cell+1]-indptr[ext_cell]
    rr = np.repeat(np.arange(ne), cnt)
    offs = np.concatenate([np.arange(indptr[c], indptr[c+1]) for c in ext_cell]) if ne < 400000 else None
    if offs is None:
        starts = np.repeat(indptr[ext_cell], cnt); within = np.arange(len(rr)) - np.repeat(np.cumsum(cnt)-cnt, cnt); offs = starts + within
    jj = indices[offs]; vv = data[offs]
    k2 = ext_tile[rr].astype(np.int64)*n + jj
    pos = np.searchsorted(key_sorted, k2); pos[pos>=ne] = ne-1
    ok = key_sorted[pos] == k2
    Aext = sp.csr_matrix((vv[ok], (rr[ok], sorter[pos[ok]])), shape=(ne, ne))
    Lo = sp.tril(Aext,0).tocsr(); Up = sp.triu(Aext,1).tocsr()
    core = tile_of[ext_cell] == ext_tile
    def apply(u):
        ue = u[ext_cell]; z = np.zeros(ne)
        for s in range(S): z = spla.spsolve_triangular(Lo, ue - Up @ z, lower=True)
        out = np.zeros(n); out[ext_cell[core]] = z[core]
        return out
    return apply, ne/n, nt

def run(apply, name):
    napp=[0]
    def Mop(u): napp[0]+=1; return apply(u)
    M = spla.LinearOperator((n,n), matvec=Mop)
    it=[0]
    x, info = spla.bicgstab(A, b, x0=x0.copy(), rtol=1e-13, atol=0.0, M=M, maxiter=200, callback=lambda xk: it.__setitem__(0, it[0]+1))
    print(f"{name:50s} iterations {it[0]:3d} applications {napp[0]:3d} relres {np.linalg.norm(b-A@x)/np.linalg.norm(b):.1e}", flush=True)

for S in (5,):
    run(make_global_gs(S), f"global GS, {S} sweeps")
for tile, depth, S in ((24,0,5),(24,4,5),(24,8,5),(24,8,10),(16,8,10),(32,8,10),(24,12,10)):
    t0=time.time(); ap, red, nt = make_tiled_gs(S, tile, depth)
    run(ap, f"tiles {tile}x{tile} (+{depth} halo), {S} sweeps, redundancy {red:.2f}, {nt} tiles")
