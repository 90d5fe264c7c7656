"""CPU experiment (scipy emulation, no GPU): chain-blocked sweep orderings (chains of 2 / 4 cells along the strongest
flow processed as one unit of a colour).  Negative result: 0.46 / 0.36 error factor per sweep against 0.31 for the
plain cell colouring -- not pursued."""
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla, sys
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[2]))
from clearwater_riverine_b200 import synthetic
from clearwater_riverine_b200.backend import order_cells
from oracle import reference_step as ref
T=24
plan = synthetic.make_plan(300, 300, T, dt=30.0, tri_fraction=0.1, dry_fraction=0.02, courant=1.5, n_exact=100_000, seed=2)
D=0.1
adv,_,_,cdiff,dt = ref.derive_coefficients(plan.face_flow, plan.edge_velocity, plan.face_x, plan.face_y, plan.f1, plan.f2, D, plan.time_seconds)
mesh = ref.HydroMesh(plan.f1, plan.f2, plan.n_face, adv, cdiff, plan.edge_velocity, plan.volume, dt, D)
n=plan.n_real
hint = plan.face_flow.mean(0)
def matrix(t):
    lhs = ref.LHS(mesh); lhs.update_values(mesh, t)
    A = lhs.to_csr(); A.sum_duplicates()
    return sp.diags(1.0/A.diagonal()) @ A
def gs_rate(A, perm, sweeps=6):
    P = sp.csr_matrix((np.ones(n), (np.arange(n), perm)), shape=(n,n))
    Ap = (P @ A @ P.T).tocsr()
    Lo = sp.tril(Ap, 0).tocsr(); Up = sp.triu(Ap, 1).tocsr()
    u = np.random.default_rng(0).random(n)
    exact = spla.spsolve(Ap.tocsc(), u)
    z = np.zeros(n); errs=[]
    for s in range(sweeps):
        z = spla.spsolve_triangular(Lo, u - Up @ z, lower=True)
        errs.append(np.linalg.norm(z-exact)/np.linalg.norm(exact))
    return (errs[-1]/errs[1])**(1/4), errs[0]
internal = plan.f2 < n
f1i, f2i, hi = plan.f1[internal], plan.f2[internal], hint[internal]
up = np.where(hi > 0, f1i, f2i); dn = np.where(hi > 0, f2i, f1i); w = np.abs(hi)
def chains(L):
    # greedy: strongest edges first; succ/pred unique; chain length capped at L
    order = np.argsort(-w)
    succ = -np.ones(n, int); pred = -np.ones(n, int)
    head = np.arange(n); length = np.ones(n, int); tail = np.arange(n)   # per chain-head bookkeeping
    chain_of = np.arange(n)
    for e in order:
        if w[e] <= 0: break
        a, b = up[e], dn[e]
        if succ[a] != -1 or pred[b] != -1: continue
        ha, hb = chain_of[a], chain_of[b]
        if ha == hb: continue
        # a must be tail of its chain, b head of its chain
        if tail[ha] != a or hb != b: continue
        if length[ha] + length[hb] > L: continue
        succ[a] = b; pred[b] = a
        # merge chain hb into ha
        x = b
        while x != -1:
            chain_of[x] = ha; x = succ[x]
        tail[ha] = tail[hb]; length[ha] += length[hb]
    return chain_of, pred, succ
A3 = matrix(3); A12 = matrix(12)
for nc in (11, 16):
    p, _, _ = order_cells(plan.f1, plan.f2, plan.n_face, True, nc, hint)
    print('cells   nc', nc, 'factor/sweep', ['%.3f'%gs_rate(A, np.argsort(p))[0] for A in (A3, A12)])
for L in (2, 4):
    chain_of, pred, succ = chains(L)
    heads = np.unique(chain_of); nq = len(heads)
    qid = -np.ones(n, int); qid[heads] = np.arange(nq)
    q = qid[chain_of]
    # position within chain
    pos = np.zeros(n, int)
    for hd in heads:
        x = hd; k = 0
        while x != -1:
            pos[x] = k; k += 1; x = succ[x]
    # quotient edges
    qa, qb = q[f1i], q[f2i]
    m = qa != qb
    key = np.minimum(qa[m], qb[m]).astype(np.int64) * nq + np.maximum(qa[m], qb[m])
    flow_signed = np.where(qa[m] < qb[m], 1.0, -1.0) * np.where(hi[m] > 0, 1.0, -1.0) * w[m]   # + : from min to max
    uk, inv = np.unique(key, return_inverse=True)
    fsum = np.bincount(inv, flow_signed)
    e1 = (uk // nq).astype(np.int32); e2 = (uk % nq).astype(np.int32)
    # order_cells needs max(f1) == nq-1: orient so that... add ghost edge from nq-1
    F1 = np.concatenate([e1, [nq-1]]).astype(np.int32); F2 = np.concatenate([e2, [nq]]).astype(np.int32)
    H = np.concatenate([fsum, [0.0]]).astype(np.float32)
    print('L', L, 'chains', nq, 'mean len %.2f'%(n/nq), 'max quotient degree', np.bincount(np.concatenate([e1,e2])).max())
    for nc in (11, 16):
        pq, cptr, nl = order_cells(F1, F2, nq+1, True, nc, H)
        ncol = len(cptr)-1
        keyc = pq[q].astype(np.int64) * 16 + pos
        perm = np.argsort(keyc, kind='stable')
        print('   chains L', L, 'nc', nc, '->', ncol, 'colours; factor/sweep', ['%.3f'%gs_rate(A, perm)[0] for A in (A3, A12)])
