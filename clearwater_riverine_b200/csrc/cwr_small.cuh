// Small-mesh solver path: one CTA per constituent / scenario runs the WHOLE preconditioned
// BiCGSTAB solve of its column inside one kernel launch.
//
// Meshes like the Ohio River model (2,943 cells) are launch-latency bound: a solve is ~100 SpMVs of
// a 140 KB matrix, and a kernel launch costs more than the SpMV it carries.  Here the iteration loop,
// its dot products (block-level, __syncthreads-based, deterministic) and the convergence test all
// live in the kernel; the matrix streams from L2, the work vectors are column-contiguous scratch that
// stays in L1/L2.  Columns are independent (they only share the matrix), so an ensemble of scenarios
// is simply gridDim.x = K CTAs with no grid-wide synchronisation, and every column stops as soon as
// it has converged.  Same arithmetic as the multi-CTA path (same preconditioner, same recurrences).
#pragma once
#include "cwr_kernels.cuh"

namespace cwr {

constexpr int kSmallThreads = 1024;

struct SmallStats {
    int max_iterations;
    int flags_or;
    int restarts;
    int not_converged;
    unsigned long long max_relres2_bits;   // max over columns/steps of ||r||^2/||b||^2 (bit pattern of a double >= 0)
    unsigned long long sum_iterations;     // over columns and steps
};

template <int ND>
__device__ __forceinline__ void cta_sum(double (&v)[ND], double* smem /* [ND][32] */) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int d = 0; d < ND; ++d) v[d] += __shfl_xor_sync(0xffffffffu, v[d], off);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();                       // smem free from the previous use
    if (lane == 0)
#pragma unroll
        for (int d = 0; d < ND; ++d) smem[d * 32 + warp] = v[d];
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            double s = lane < (int)(blockDim.x >> 5) ? smem[d * 32 + lane] : 0.0;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if (lane == 0) smem[d * 32] = s;
        }
    }
    __syncthreads();
#pragma unroll
    for (int d = 0; d < ND; ++d) v[d] = smem[d * 32];
}

// (L z)_i for a column-contiguous z
__device__ __forceinline__ double ell_row_dot(const DeviceModel& M, int i, const double* z) {
    double s = 0.0;
    const int W = M.W;
    for (int w = 0; w < W; w += 4) {
        int4 c4 = __ldg(reinterpret_cast<const int4*>(M.ell_col + (size_t)i * W + w));
        c4.x &= kColMask; c4.y &= kColMask; c4.z &= kColMask; c4.w &= kColMask;
        const double2 v01 = *reinterpret_cast<const double2*>(M.val + (size_t)i * W + w);
        const double2 v23 = *reinterpret_cast<const double2*>(M.val + (size_t)i * W + w + 2);
        s = fma(v23.y, z[c4.w], fma(v23.x, z[c4.z], fma(v01.y, z[c4.y], fma(v01.x, z[c4.x], s))));
    }
    return s;
}

// z = (I + N + ... + N^(m-1)) u  with N = -L; returns the buffer holding z
__device__ __forceinline__ const double* small_precondition(const DeviceModel& M, int n, int m_steps, const double* u,
                                                            double* dst, double* other) {
    const int J = m_steps - 1;
    const double* z = u;
    for (int j = 1; j <= J; ++j) {
        double* out = ((J - j) % 2 == 0) ? dst : other;
        for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = u[i] - ell_row_dot(M, i, z);
        __syncthreads();
        z = out;
    }
    return z;
}

__global__ void __launch_bounds__(kSmallThreads, 1) k_solve_small(DeviceModel M, int m_steps, SmallStats* stats) {
    __shared__ double red[4 * 32];
    pdl_enter();
    const int k = blockIdx.x, n = M.n, K = M.K;
    double* __restrict__ x = M.sp->state_t1 + k;          // stride K
    const double* __restrict__ b = M.b + k;               // stride K
    const size_t col0 = (size_t)k * n;
    double *r = M.r + col0, *rhat = M.rhat + col0, *p = M.p + col0, *v = M.v + col0, *t = M.tt + col0;
    double *phb = (double*)M.ph + col0, *shb = (double*)M.sh + col0, *tmp = (double*)M.tmp + col0, *xc = M.xc + col0;
    int flags = 0, iters = 0, restarts = 0;
    double bb = 0.0, rr = 0.0;

    for (int i = threadIdx.x; i < n; i += blockDim.x) xc[i] = x[(size_t)i * K];   // contiguous copy of the iterate
    __syncthreads();

    for (;;) {
        // r = b - A x ; rhat = p = r
        double d2[2] = {0.0, 0.0};
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const double bi = b[(size_t)i * K];
            const double ri = bi - (xc[i] + ell_row_dot(M, i, xc));
            r[i] = ri; rhat[i] = ri; p[i] = ri;
            d2[0] = fma(ri, ri, d2[0]); d2[1] = fma(bi, bi, d2[1]);
        }
        cta_sum<2>(d2, red);
        rr = d2[0]; bb = d2[1];
        if (!(rr == rr) || !(bb == bb) || isinf(rr) || isinf(bb)) {
            flags |= FL_NAN;
            for (int i = threadIdx.x; i < n; i += blockDim.x) xc[i] = qnan();
            break;
        }
        if (bb == 0.0 && rr != 0.0) {                    // b == 0  =>  x = 0
            flags |= FL_ZERO_RHS | FL_CONVERGED;
            for (int i = threadIdx.x; i < n; i += blockDim.x) xc[i] = 0.0;
            rr = 0.0;
            break;
        }
        if (rr <= M.tol2 * bb) { flags |= FL_CONVERGED; break; }
        double rho = rr;
        bool breakdown = false;
        while (iters < M.max_iter) {
            const double* ph = small_precondition(M, n, m_steps, p, phb, tmp);
            double d1[1] = {0.0};
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const double y = ph[i] + ell_row_dot(M, i, ph);
                v[i] = y;
                d1[0] = fma(rhat[i], y, d1[0]);
            }
            cta_sum<1>(d1, red);
            if (d1[0] == 0.0 || !(d1[0] == d1[0])) { breakdown = true; break; }
            const double alpha = rho / d1[0];
            double dh[1] = {0.0};
            for (int i = threadIdx.x; i < n; i += blockDim.x) { r[i] = fma(-alpha, v[i], r[i]); dh[0] = fma(r[i], r[i], dh[0]); }     // s
            cta_sum<1>(dh, red);
            if (dh[0] <= M.tol2 * bb) {                      // converged at the half step: x += alpha p^
                for (int i = threadIdx.x; i < n; i += blockDim.x) xc[i] = fma(alpha, ph[i], xc[i]);
                rr = dh[0]; ++iters; flags |= FL_CONVERGED;
                break;
            }
            const double* sh = small_precondition(M, n, m_steps, r, shb, tmp);
            double d4[4] = {0.0, 0.0, 0.0, 0.0};
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const double y = sh[i] + ell_row_dot(M, i, sh);
                t[i] = y;
                const double si = r[i], rh = rhat[i];
                d4[0] = fma(y, si, d4[0]); d4[1] = fma(y, y, d4[1]); d4[2] = fma(rh, y, d4[2]); d4[3] = fma(rh, si, d4[3]);
            }
            cta_sum<4>(d4, red);
            const double omega = d4[1] > 0.0 ? d4[0] / d4[1] : 0.0;
            const double rho_new = d4[3] - omega * d4[2];
            double beta = 0.0;
            bool stagnated = false;
            if (omega != 0.0 && rho != 0.0) beta = (rho_new / rho) * (alpha / omega);
            else if (d4[1] > 0.0) stagnated = true;
            if (!(beta == beta) || isinf(beta)) { beta = 0.0; stagnated = true; }
            double dr[1] = {0.0};
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const double si = r[i], pv = p[i];
                xc[i] = fma(omega, sh[i], fma(alpha, ph[i], xc[i]));
                const double rn = fma(-omega, t[i], si);
                r[i] = rn;
                p[i] = fma(beta, fma(-omega, v[i], pv), rn);
                dr[0] = fma(rn, rn, dr[0]);
            }
            cta_sum<1>(dr, red);
            rr = dr[0];
            ++iters;
            rho = rho_new;
            if (!(rr == rr) || isinf(rr)) { flags |= FL_NAN; break; }
            if (rr <= M.tol2 * bb) { flags |= FL_CONVERGED; break; }
            if (stagnated) { breakdown = true; break; }
        }
        if ((flags & (FL_CONVERGED | FL_NAN)) || iters >= M.max_iter) break;
        if (breakdown && restarts < 3) { ++restarts; continue; }     // new shadow residual from the current iterate
        if (breakdown) flags |= FL_BREAKDOWN;
        break;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) x[(size_t)i * K] = xc[i];
    if (threadIdx.x == 0) {
        M.colflags[k] = flags; M.coliters[k] = iters;
        M.sc[SC_BNORM2 * K + k] = bb; M.sc[SC_RNORM2 * K + k] = rr;
        atomicMax(&stats->max_iterations, iters);
        atomicAdd(&stats->sum_iterations, (unsigned long long)iters);
        atomicOr(&stats->flags_or, flags & (FL_BREAKDOWN | FL_NAN));
        atomicMax(&stats->restarts, restarts);
        if (!(flags & FL_CONVERGED)) atomicAdd(&stats->not_converged, 1);
        const double rel2 = bb > 0.0 ? rr / bb : (rr > 0.0 ? __longlong_as_double(0x7ff0000000000000LL) : 0.0);
        if (rel2 == rel2) atomicMax(&stats->max_relres2_bits, (unsigned long long)__double_as_longlong(rel2));
    }
}



// ---------------------------------------------------------------------------------------------
// Tiny meshes (the Ohio River model: 2,943 cells): the whole solve of one column ON CHIP.
// One CTA per constituent / scenario; the matrix (column-major ELL: fp64 values + 16-bit column
// indices), the gathered vector z and the preconditioner's input u live in shared memory (<= 227 KB),
// the BiCGSTAB vectors (x, r, p, v, t, p^) in registers -- a thread owns rows tid, tid + 512, ... --
// and the shadow residual in a column-contiguous scratch only its owner thread touches.  The
// preconditioner is the same flow-aligned multicolour Gauss-Seidel as on large meshes with a CTA barrier
// between colours.
//
// A solve is ~250 barrier-separated phases (information travels one flow level per colour step, and
// ~100 levels matter at Courant 2.5), so the cost of a PHASE is what counts, not bytes:
//  * a colour's rows (a contiguous range, ~n / colours of them) are taken by the first warps of the CTA,
//    row = first row of the colour + thread id; the other warps go straight to the barrier.  (The first
//    version walked every thread's own rows with a colour predicate: 32 warps x ~60 instructions per
//    phase, issue-bound at ~900 cycles per phase.)
//  * 512 threads: half the warps at every barrier;
//  * dot products cost ONE barrier: per-warp partials into a double-buffered array, every warp adds the
//    partials itself in the same order (bitwise identical totals in all warps).
// Same recurrences, stopping test, restart and NaN / zero-rhs rules as k_solve_small.
// ---------------------------------------------------------------------------------------------
constexpr int kTinyThreads = 512;
constexpr int kTinyMaxRows = 8;          // rows per thread: n <= 4096 (and the shared-memory limit)
constexpr int kTinyStaticSmem = 4096;    // red[] + colour pointers, rounded up

__host__ __device__ inline size_t tiny_smem_bytes(int n, int W) {
    // val (n*W f64) | z, u (n f64 each) | idx (n*W u16)
    return (size_t)n * W * 8 + (size_t)2 * n * 8 + (((size_t)n * W * 2 + 15) & ~(size_t)15);
}

// one barrier: totals of ND partial sums, identical bits in every thread; `red` is [2][ND][32], `parity` alternates
template <int ND, int NWARPS = kTinyThreads / 32>
__device__ __forceinline__ void tiny_sum(double (&v)[ND], double* red, int& parity) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int d = 0; d < ND; ++d) v[d] += __shfl_xor_sync(0xffffffffu, v[d], off);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* buf = red + parity * (4 * 32);
    parity ^= 1;
    if (lane == 0)
#pragma unroll
        for (int d = 0; d < ND; ++d) buf[d * 32 + warp] = v[d];
    __syncthreads();
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        double s = lane < NWARPS ? buf[d * 32 + lane] : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        v[d] = s;
    }
}

// W4: ELL width 4 (quad / triangle meshes): a row's four values are two 16-byte shared loads, its four
// 16-bit column indices one 8-byte load.
template <int RPT, bool W4>
__global__ void __launch_bounds__(kTinyThreads, 1) k_solve_tiny(DeviceModel M, int n_sweeps, SmallStats* stats) {
    extern __shared__ double tiny_smem[];
    __shared__ double red[2 * 4 * 32];
    __shared__ int cptr[65];
    pdl_enter();
    const int k = blockIdx.x, n = M.n, K = M.K, W = M.W, nc = M.n_colors;
    double* sval = tiny_smem;                       // [W][n]   (W4: double2 [2][n])
    double* z = sval + (size_t)n * W;               // [n] the gathered vector
    double* su = z + n;                             // [n] what the sweeps are applied to
    unsigned short* sidx = reinterpret_cast<unsigned short*>(su + n);   // [W][n] (W4: [n][4]), bit 15 = visited later in a sweep
    double* __restrict__ xg = M.sp->state_t1 + k;   // stride K
    const double* __restrict__ bg = M.b + k;
    double* __restrict__ rhat = M.rhat + (size_t)k * n;   // element i only ever touched by the thread that owns row i
    int parity = 0;

    // ---- matrix -> shared memory (column-major) ------------------------------------------------------
    for (int q = threadIdx.x; q < n * W; q += kTinyThreads) {
        const int i = q / W, w = q % W;
        const int32_t cj = M.ell_col[q];
        const unsigned short packed = (unsigned short)((cj & 0x7fff) | (cj < 0 ? 0x8000 : 0));
        if (W4) { sval[((size_t)(w >> 1) * n + i) * 2 + (w & 1)] = M.val[q]; sidx[(size_t)i * 4 + w] = packed; }
        else { sval[(size_t)w * n + i] = M.val[q]; sidx[(size_t)w * n + i] = packed; }
    }
    if (threadIdx.x <= nc) cptr[threadIdx.x] = M.color_ptr[threadIdx.x];
    int row[RPT]; bool on[RPT];
    double x[RPT], r[RPT], p[RPT], v[RPT], t[RPT], ph[RPT];
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
        row[q] = threadIdx.x + q * kTinyThreads; on[q] = row[q] < n;
        x[q] = on[q] ? xg[(size_t)row[q] * K] : 0.0;
        r[q] = p[q] = v[q] = t[q] = ph[q] = 0.0;
    }
    __syncthreads();

    auto row_dot = [&](int i, bool first) {      // (L z)_i ; first: skip neighbours not visited yet (z = 0 there)
        if (W4) {
            const uint2 id = reinterpret_cast<const uint2*>(sidx)[i];
            const double2 a = reinterpret_cast<const double2*>(sval)[i], b2 = reinterpret_cast<const double2*>(sval)[n + i];
            const unsigned j0 = id.x & 0xffffu, j1 = id.x >> 16, j2 = id.y & 0xffffu, j3 = id.y >> 16;
            const double z0 = (first && (j0 & 0x8000u)) ? 0.0 : z[j0 & 0x7fffu], z1 = (first && (j1 & 0x8000u)) ? 0.0 : z[j1 & 0x7fffu];
            const double z2 = (first && (j2 & 0x8000u)) ? 0.0 : z[j2 & 0x7fffu], z3 = (first && (j3 & 0x8000u)) ? 0.0 : z[j3 & 0x7fffu];
            return fma(b2.y, z3, fma(b2.x, z2, fma(a.y, z1, a.x * z0)));
        }
        double s = 0.0;
        for (int w = 0; w < W; ++w) {
            const unsigned short cj = sidx[(size_t)w * n + i];
            if (first && (cj & 0x8000)) continue;
            s = fma(sval[(size_t)w * n + i], z[cj & 0x7fff], s);
        }
        return s;
    };
    // z = M^-1 su (su complete and synchronised): n_sweeps multicolour Gauss-Seidel sweeps from z = 0; ends synchronised
    auto precondition = [&]() {
        if (n_sweeps <= 0) {
#pragma unroll
            for (int q = 0; q < RPT; ++q) if (on[q]) z[row[q]] = su[row[q]];
            __syncthreads();
            return;
        }
        for (int s = 0; s < n_sweeps; ++s)
            for (int c = 0; c < nc; ++c) {
                const int hi = cptr[c + 1];
                for (int i = cptr[c] + (int)threadIdx.x; i < hi; i += kTinyThreads) z[i] = su[i] - row_dot(i, s == 0);
                __syncthreads();
            }
    };
    // y = A z for this thread's rows (z complete and synchronised)
    auto product = [&](double (&y)[RPT]) {
#pragma unroll
        for (int q = 0; q < RPT; ++q) y[q] = on[q] ? z[row[q]] + row_dot(row[q], false) : 0.0;
    };

    int flags = 0, iters = 0, restarts = 0;
    double bb = 0.0, rr = 0.0;
    for (;;) {
        // r = b - A x ; rhat = p = r
#pragma unroll
        for (int q = 0; q < RPT; ++q) if (on[q]) z[row[q]] = x[q];
        __syncthreads();
        double ax[RPT];
        product(ax);
        double d2[2] = {0.0, 0.0};
#pragma unroll
        for (int q = 0; q < RPT; ++q)
            if (on[q]) {
                const double bi = bg[(size_t)row[q] * K];
                r[q] = bi - ax[q]; p[q] = r[q]; rhat[row[q]] = r[q]; su[row[q]] = r[q];
                d2[0] = fma(r[q], r[q], d2[0]); d2[1] = fma(bi, bi, d2[1]);
            }
        tiny_sum<2>(d2, red, parity);                    // its barrier also publishes su = p and retires the reads of z
        rr = d2[0]; bb = d2[1];
        if (!(rr == rr) || !(bb == bb) || isinf(rr) || isinf(bb)) {
            flags |= FL_NAN;
#pragma unroll
            for (int q = 0; q < RPT; ++q) x[q] = qnan();
            break;
        }
        if (bb == 0.0 && rr != 0.0) {                    // b == 0  =>  x = 0
            flags |= FL_ZERO_RHS | FL_CONVERGED;
#pragma unroll
            for (int q = 0; q < RPT; ++q) x[q] = 0.0;
            rr = 0.0;
            break;
        }
        if (rr <= M.tol2 * bb) { flags |= FL_CONVERGED; break; }
        double rho = rr;
        bool breakdown = false;
        while (iters < M.max_iter) {
            precondition();                              // z = p^
#pragma unroll
            for (int q = 0; q < RPT; ++q) ph[q] = on[q] ? z[row[q]] : 0.0;
            product(v);
            double d1[1] = {0.0};
#pragma unroll
            for (int q = 0; q < RPT; ++q) if (on[q]) d1[0] = fma(rhat[row[q]], v[q], d1[0]);
            tiny_sum<1>(d1, red, parity);                // (all reads of z and su are behind this barrier)
            if (d1[0] == 0.0 || !(d1[0] == d1[0])) { breakdown = true; break; }
            const double alpha = rho / d1[0];
            double dh[1] = {0.0};
#pragma unroll
            for (int q = 0; q < RPT; ++q) {
                r[q] = fma(-alpha, v[q], r[q]);                                                                // s
                if (on[q]) { dh[0] = fma(r[q], r[q], dh[0]); su[row[q]] = r[q]; }
            }
            tiny_sum<1>(dh, red, parity);                // publishes su = s
            if (dh[0] <= M.tol2 * bb) {                      // converged at the half step: x += alpha p^
#pragma unroll
                for (int q = 0; q < RPT; ++q) x[q] = fma(alpha, ph[q], x[q]);
                rr = dh[0]; ++iters; flags |= FL_CONVERGED;
                break;
            }
            precondition();                              // z = s^
            product(t);
            double d4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int q = 0; q < RPT; ++q)
                if (on[q]) {
                    const double rh = rhat[row[q]];
                    d4[0] = fma(t[q], r[q], d4[0]); d4[1] = fma(t[q], t[q], d4[1]);
                    d4[2] = fma(rh, t[q], d4[2]); d4[3] = fma(rh, r[q], d4[3]);
                }
            double shq[RPT];                             // s^ of the own rows, read before the barrier that frees z
#pragma unroll
            for (int q = 0; q < RPT; ++q) shq[q] = on[q] ? z[row[q]] : 0.0;
            tiny_sum<4>(d4, red, parity);
            const double omega = d4[1] > 0.0 ? d4[0] / d4[1] : 0.0;
            const double rho_new = d4[3] - omega * d4[2];
            double beta = 0.0;
            bool stagnated = false;
            if (omega != 0.0 && rho != 0.0) beta = (rho_new / rho) * (alpha / omega);
            else if (d4[1] > 0.0) stagnated = true;
            if (!(beta == beta) || isinf(beta)) { beta = 0.0; stagnated = true; }
            double dr[1] = {0.0};
#pragma unroll
            for (int q = 0; q < RPT; ++q)
                if (on[q]) {
                    x[q] = fma(omega, shq[q], fma(alpha, ph[q], x[q]));
                    const double rn = fma(-omega, t[q], r[q]);
                    r[q] = rn;
                    p[q] = fma(beta, fma(-omega, v[q], p[q]), rn);
                    su[row[q]] = p[q];
                    dr[0] = fma(rn, rn, dr[0]);
                }
            tiny_sum<1>(dr, red, parity);                // publishes su = p
            rr = dr[0];
            ++iters;
            rho = rho_new;
            if (!(rr == rr) || isinf(rr)) { flags |= FL_NAN; break; }
            if (rr <= M.tol2 * bb) { flags |= FL_CONVERGED; break; }
            if (stagnated) { breakdown = true; break; }
        }
        if ((flags & (FL_CONVERGED | FL_NAN)) || iters >= M.max_iter) break;
        if (breakdown && restarts < 3) { ++restarts; __syncthreads(); continue; }     // new shadow residual from the current iterate
        if (breakdown) flags |= FL_BREAKDOWN;
        break;
    }
#pragma unroll
    for (int q = 0; q < RPT; ++q) if (on[q]) xg[(size_t)row[q] * K] = x[q];
    if (threadIdx.x == 0) {
        M.colflags[k] = flags; M.coliters[k] = iters;
        M.sc[SC_BNORM2 * K + k] = bb; M.sc[SC_RNORM2 * K + k] = rr;
        atomicMax(&stats->max_iterations, iters);
        atomicAdd(&stats->sum_iterations, (unsigned long long)iters);
        atomicOr(&stats->flags_or, flags & (FL_BREAKDOWN | FL_NAN));
        atomicMax(&stats->restarts, restarts);
        if (!(flags & FL_CONVERGED)) atomicAdd(&stats->not_converged, 1);
        const double rel2 = bb > 0.0 ? rr / bb : (rr > 0.0 ? __longlong_as_double(0x7ff0000000000000LL) : 0.0);
        if (rel2 == rel2) atomicMax(&stats->max_relres2_bits, (unsigned long long)__double_as_longlong(rel2));
    }
}


// ---------------------------------------------------------------------------------------------
// k_solve_chip: the on-chip solve again, organised around the COLOUR STEP.
//
// ncu of k_solve_tiny on the Ohio-shaped mesh (profiles/r02tiny1_*): 240 colour steps + ~20 vector phases
// in 194 k cycles = ~800 cycles per colour step, and the step is one dependent chain of ~65 instructions
// (colour pointer -> row -> index word -> unpack -> gathers -> 4 chained DFMA -> store -> barrier) issued at one
// instruction per ~11 cycles; bytes and flops are irrelevant.  So this kernel takes everything that does not depend
// on z out of the chain:
//  * thread t owns row (first row of colour c) + t of EVERY colour c -- the rows it sweeps are the rows whose
//    BiCGSTAB vectors it keeps: the right-hand side of a sweep (p or r) is a register, nothing is published for it;
//  * its rows' matrix values (4 x fp64) and neighbour indices (4 x 16 bit) stay in REGISTERS for the whole solve
//    (10 per colour); a colour step loads nothing but the four gathered z (shared-memory wavefronts are the
//    second limit after latency: 8 warps x 4 x 64-bit gathers ~ 90 per step);
//  * the vectors a step does not need (x, p^, v, t, r^) live in private shared-memory columns [colour][thread]
//    (conflict-free), touched only in the vector phases;
//  * the four products are summed as a tree (two DFMA chains of depth 2 + one add);
//  * z starts at zero (zeroed by its owners behind a barrier that is there anyway), so the first sweep
//    needs no "not visited yet" predicates;
//  * 256 threads = the rows of a colour (n / colours <= 256): 8 warps at every barrier;
//  * the global loads of the set-up (matrix, x, b: ~40 per thread) are issued in batches, not one per dependent
//    use (the first version spent 27 % of its cycles there).
// Needs ELL width 4, <= NS colours of <= 256 rows; anything else takes k_solve_tiny.  Same recurrences,
// stopping test, restart and NaN / zero-rhs rules as k_solve_small.
// ---------------------------------------------------------------------------------------------
constexpr int kChipThreads = 256;

__host__ __device__ inline size_t chip_smem_bytes(int n, int NS) {
    // x, p^, v, t, r^, p [NS][256] f64 | z [n + 1] f64 | zf [n + 2] f32 (p and zf: only the kernel with fp32 sweeps uses them;
    // element n of z / zf takes the stores of threads without a row in the colour)
    return (size_t)NS * kChipThreads * 48 + (size_t)(n + 1) * 8 + (size_t)(n + 2) * 4;
}

// F32: the sweeps run in fp32 (options.precond_precision = 32, the default): the gathered vector is 4 bytes per row
// (half the shared-memory wavefronts of a colour step) and the chain of a step is FMUL/FFMA/FADD.  What the sweeps are
// given (p or r) is scaled by a power of two near 1 / ||r|| first, so the fp32 range never matters; p^ / s^ = the fp32
// result widened exactly, and v = A p^, t = A s^, every dot product and the recurrences stay fp64 with the fp64 matrix
// values: the converged answer is the fp64 one, as on the large path.
// The registers then hold fp32 copies of the matrix values (and p moves to a private shared-memory column); the fp64 values
// are re-read from global memory (L2) in the few vector phases that need them (v = A p^, t = A s^, r = b - A x).
// Measured on the Ohio-shaped mesh (profiles/r02chip1_*): fp64 sweeps 0.071 ms per step, fp32 0.069 (64 scenarios 0.080 / 0.078),
// same iteration counts; keeping the fp64 values in registers and narrowing them inside the colour step (4 F2F per step, off
// the dependent chain) was slower than fp64 (0.073) and is gone.
// FULL: the mesh has exactly NS colours (no "is there a colour c" test in the colour step).  A colour step has no branch:
// threads without a row in the colour compute on row 0's neighbours with zero values and store into a spare element
// (ncu, profiles/r02chip2_*: the two test-and-branch pairs at the head of a step held 40 % of its stall samples).
template <int NS, bool F32, bool FULL>
__global__ void __launch_bounds__(kChipThreads, 1) k_solve_chip(DeviceModel M, int n_sweeps, SmallStats* stats) {
    constexpr bool REG64 = !F32;                 // fp64 matrix values in registers
    extern __shared__ double chip_smem[];
    __shared__ double red[2 * 4 * 32];
    __shared__ int cptr[NS + 1];
    constexpr int NT = kChipThreads, NW = kChipThreads / 32;
    pdl_launch_dependents();
    const int tid = threadIdx.x, k = blockIdx.x, n = M.n, K = M.K, nc = M.n_colors;
    double* sx = chip_smem + tid;                               // [NS][NT] private columns: element c at [c * NT]
    double* sph = sx + NS * NT;
    double* sv = sph + NS * NT;
    double* st = sv + NS * NT;
    double* srh = st + NS * NT;
    double* spp = srh + NS * NT;                                // (fp32 sweeps: p lives here, its registers go to the compiler)
    double* z = chip_smem + 6 * NS * NT;                        // [n] the gathered vector
    float* zf = reinterpret_cast<float*>(z + n + 1);            // [n] ... of the fp32 sweeps
    const double* __restrict__ bg = M.b + k;
    int parity = 0;

    // (what follows down to pdl_wait() reads the mesh only -- colour pointers and neighbour indices, written at set-up: it
    // runs while the kernels that build this step's matrix and right-hand side are still at work)
    if (tid <= NS) cptr[tid] = M.color_ptr[tid < nc ? tid : nc];
    __syncthreads();
#define CHIP_ROW(c) (cptr[c] + tid)
#define CHIP_ON(c) ((on >> (c)) & 1u)
    unsigned on = 0;
#pragma unroll
    for (int c = 0; c < NS; ++c) if (c < nc && CHIP_ROW(c) < cptr[c + 1]) on |= 1u << c;
    double2 va[REG64 ? NS : 1], vb[REG64 ? NS : 1]; uint2 nb[NS];
    float4 vf[REG64 ? 1 : NS];
    double r[NS], p[F32 ? 1 : NS];
    auto getp = [&](int c) -> double { if constexpr (F32) return spp[c * NT]; else return p[c]; };
    auto setp = [&](int c, double v) { if constexpr (F32) spp[c * NT] = v; else p[c] = v; };
    // ---- the thread's rows: neighbour indices -> registers
#pragma unroll
    for (int c0 = 0; c0 < NS; c0 += 4) {
        int4 cj[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (c0 + q < NS) cj[q] = __ldg(reinterpret_cast<const int4*>(M.ell_col + (size_t)(CHIP_ON(c0 + q) ? CHIP_ROW(c0 + q) : 0) * 4));
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = c0 + q;
            if (c < NS) {
                nb[c].x = (unsigned)(cj[q].x & 0x7fff) | ((unsigned)(cj[q].y & 0x7fff) << 16);
                nb[c].y = (unsigned)(cj[q].z & 0x7fff) | ((unsigned)(cj[q].w & 0x7fff) << 16);
                if (!CHIP_ON(c)) nb[c] = make_uint2(0u, 0u);
            }
        }
    }
    pdl_wait();
    double* __restrict__ xg = M.sp->state_t1 + k;               // stride K
    // ---- matrix values -> registers, x -> its column (loads issued four rows at a time)
#pragma unroll
    for (int c0 = 0; c0 < NS; c0 += 4) {
        double xq[4]; double2 ta[4], tb[4]; float4 tf[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = c0 + q;
            if (c < NS) {
                const size_t row = CHIP_ON(c) ? CHIP_ROW(c) : 0;
                if constexpr (REG64) {
                    ta[q] = __ldg(reinterpret_cast<const double2*>(M.val + row * 4));
                    tb[q] = __ldg(reinterpret_cast<const double2*>(M.val + row * 4 + 2));
                } else tf[q] = __ldg(reinterpret_cast<const float4*>(M.valf + row * 4));     // (k_assemble's fp32 copy: same rounding)
                xq[q] = xg[row * K];
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = c0 + q;
            if (c < NS) {
                if (!CHIP_ON(c)) { ta[q] = make_double2(0.0, 0.0); tb[q] = ta[q]; tf[q] = make_float4(0.f, 0.f, 0.f, 0.f); xq[q] = 0.0; }
                if constexpr (REG64) { va[c] = ta[q]; vb[c] = tb[q]; }
                else vf[c] = tf[q];
                sx[c * NT] = xq[q];
                sph[c * NT] = 0.0; sv[c * NT] = 0.0; st[c * NT] = 0.0; srh[c * NT] = 0.0;
                r[c] = 0.0; setp(c, 0.0);
            }
        }
    }

    // the fp64 matrix values of the thread's row of colour c (rows that are off: whatever row 0 holds, the callers discard it)
    auto mat = [&](int c, double2& a, double2& b) {
        if constexpr (REG64) { a = va[c]; b = vb[c]; }
        else {
            const size_t row = CHIP_ON(c) ? CHIP_ROW(c) : 0;
            a = __ldg(reinterpret_cast<const double2*>(M.val + row * 4));
            b = __ldg(reinterpret_cast<const double2*>(M.val + row * 4 + 2));
        }
    };
    // the vector phases walk the colours four at a time: the matrix values of four rows are requested together (with
    // fp32 sweeps they come from L2; one colour at a time the phase paid a round trip per colour, 7 of the solve's 67 us)
    auto batched = [&](auto body) {
#pragma unroll
        for (int c0 = 0; c0 < NS; c0 += 4) {
            double2 a[4], b[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) if (c0 + q < NS) mat(c0 + q, a[q], b[q]);
#pragma unroll
            for (int q = 0; q < 4; ++q) if (c0 + q < NS) body(c0 + q, a[q], b[q]);
        }
    };
    // (L z) of the thread's row of colour c, fp64 vector
    auto dot4 = [&](int c, const double2& a, const double2& b) {
        const double z0 = z[nb[c].x & 0xffffu], z1 = z[nb[c].x >> 16], z2 = z[nb[c].y & 0xffffu], z3 = z[nb[c].y >> 16];
        return fma(a.y, z1, a.x * z0) + fma(b.y, z3, b.x * z2);
    };
    // the preconditioned vector after the sweeps: own row / (L .) of the row, in fp64 whatever the sweeps ran in
    double unscale = 1.0;                       // fp32 sweeps: 2^h, undoes the scaling of what the sweeps were given
    auto zrow = [&](int c) -> double {
        if constexpr (F32) return (double)zf[CHIP_ROW(c)] * unscale;
        else return z[CHIP_ROW(c)];
    };
    auto dot4z = [&](int c, const double2& a, const double2& b) -> double {
        if constexpr (F32) {
            const double z0 = (double)zf[nb[c].x & 0xffffu], z1 = (double)zf[nb[c].x >> 16];
            const double z2 = (double)zf[nb[c].y & 0xffffu], z3 = (double)zf[nb[c].y >> 16];
            return (fma(a.y, z1, a.x * z0) + fma(b.y, z3, b.x * z2)) * unscale;    // (power of two: exact)
        } else return dot4(c, a, b);
    };
    // z = M^-1 u: n_sweeps multicolour Gauss-Seidel sweeps; z is ZERO and synchronised on entry, complete and synchronised on exit
    // (norm2: a squared norm of the size of u's, for the fp32 scaling)
    auto precondition = [&](auto u, double norm2) {
        if constexpr (F32) {
            // scale = 2^-h with h = half the exponent of norm2: u * scale is O(1) whatever the units of the system
            const int e = ((__double2hiint(norm2) >> 20) & 0x7ff) - 1023;
            const int hh = e / 2;
            const double scale = __hiloint2double((1023 - hh) << 20, 0);
            unscale = __hiloint2double((1023 + hh) << 20, 0);
            float uf[NS];
#pragma unroll
            for (int c = 0; c < NS; ++c) uf[c] = (float)(u(c) * scale);
            if (n_sweeps <= 0) {
#pragma unroll
                for (int c = 0; c < NS; ++c) if (CHIP_ON(c)) zf[CHIP_ROW(c)] = uf[c];
                __syncthreads();
                return;
            }
            for (int s = 0; s < n_sweeps; ++s) {
#pragma unroll
                for (int c = 0; c < NS; ++c) {
                    if (FULL || c < nc) {
                        const float a0 = vf[c].x, a1 = vf[c].y, a2 = vf[c].z, a3 = vf[c].w;
                        const float z0 = zf[nb[c].x & 0xffffu], z1 = zf[nb[c].x >> 16], z2 = zf[nb[c].y & 0xffffu], z3 = zf[nb[c].y >> 16];
                        const float zi = uf[c] - (fmaf(a1, z1, a0 * z0) + fmaf(a3, z3, a2 * z2));
                        zf[CHIP_ON(c) ? CHIP_ROW(c) : n] = zi;
                        __syncthreads();
                    }
                }
            }
        } else {
            if (n_sweeps <= 0) {
#pragma unroll
                for (int c = 0; c < NS; ++c) if (CHIP_ON(c)) z[CHIP_ROW(c)] = u(c);
                __syncthreads();
                return;
            }
            for (int s = 0; s < n_sweeps; ++s) {
#pragma unroll
                for (int c = 0; c < NS; ++c) {
                    if (FULL || c < nc) {
                        const double zi = u(c) - dot4(c, va[c], vb[c]);
                        z[CHIP_ON(c) ? CHIP_ROW(c) : n] = zi;
                        __syncthreads();
                    }
                }
            }
        }
    };
    auto zero_z = [&]() {
#pragma unroll
        for (int c = 0; c < NS; ++c)
            if (CHIP_ON(c)) { if constexpr (F32) zf[CHIP_ROW(c)] = 0.f; else z[CHIP_ROW(c)] = 0.0; }
    };

    int flags = 0, iters = 0, restarts = 0;
    double bb = 0.0, rr = 0.0;
    for (;;) {
        // r = b - A x ; rhat = p = r
        double d2[2] = {0.0, 0.0};
        {
            double bq[NS];
#pragma unroll
            for (int c = 0; c < NS; ++c) bq[c] = bg[(size_t)(CHIP_ON(c) ? CHIP_ROW(c) : 0) * K];
#pragma unroll
            for (int c = 0; c < NS; ++c) if (CHIP_ON(c)) z[CHIP_ROW(c)] = sx[c * NT];
            __syncthreads();
            batched([&](int c, const double2& a, const double2& b) {
                const double bi = CHIP_ON(c) ? bq[c] : 0.0;
                const double ri = CHIP_ON(c) ? bi - (z[CHIP_ROW(c)] + dot4(c, a, b)) : 0.0;
                r[c] = ri; setp(c, ri); srh[c * NT] = ri;
                d2[0] = fma(ri, ri, d2[0]); d2[1] = fma(bi, bi, d2[1]);
            });
        }
        tiny_sum<2, NW>(d2, red, parity);                // (every read of z is behind this barrier)
        rr = d2[0]; bb = d2[1];
        if (!(rr == rr) || !(bb == bb) || isinf(rr) || isinf(bb)) {
            flags |= FL_NAN;
#pragma unroll
            for (int c = 0; c < NS; ++c) sx[c * NT] = qnan();
            break;
        }
        if (bb == 0.0 && rr != 0.0) {                    // b == 0  =>  x = 0
            flags |= FL_ZERO_RHS | FL_CONVERGED;
#pragma unroll
            for (int c = 0; c < NS; ++c) sx[c * NT] = 0.0;
            rr = 0.0;
            break;
        }
        if (rr <= M.tol2 * bb) { flags |= FL_CONVERGED; break; }
        zero_z();
        __syncthreads();
        double rho = rr;
        bool breakdown = false;
        while (iters < M.max_iter) {
            precondition(getp, rr);                      // z = p^
            double d1[1] = {0.0};
            batched([&](int c, const double2& a, const double2& b) {
                const double zi = CHIP_ON(c) ? zrow(c) : 0.0;
                const double vi = CHIP_ON(c) ? zi + dot4z(c, a, b) : 0.0;
                sph[c * NT] = zi; sv[c * NT] = vi;
                d1[0] = fma(srh[c * NT], vi, d1[0]);
            });
            tiny_sum<1, NW>(d1, red, parity);            // (all reads of z are behind this barrier)
            if (d1[0] == 0.0 || !(d1[0] == d1[0])) { breakdown = true; break; }
            const double alpha = rho / d1[0];
            double dh[1] = {0.0};
#pragma unroll
            for (int c = 0; c < NS; ++c) {
                r[c] = fma(-alpha, sv[c * NT], r[c]);                                                          // s
                dh[0] = fma(r[c], r[c], dh[0]);
            }
            zero_z();
            tiny_sum<1, NW>(dh, red, parity);            // publishes z = 0
            if (dh[0] <= M.tol2 * bb) {                      // converged at the half step: x += alpha p^
#pragma unroll
                for (int c = 0; c < NS; ++c) sx[c * NT] = fma(alpha, sph[c * NT], sx[c * NT]);
                rr = dh[0]; ++iters; flags |= FL_CONVERGED;
                break;
            }
            precondition([&](int c) { return r[c]; }, dh[0]);        // z = s^
            double d4[4] = {0.0, 0.0, 0.0, 0.0};
            batched([&](int c, const double2& a, const double2& b) {
                const double zi = CHIP_ON(c) ? zrow(c) : 0.0;                 // s^ of the own row
                const double ti = CHIP_ON(c) ? zi + dot4z(c, a, b) : 0.0;
                st[c * NT] = ti;
                sx[c * NT] = fma(alpha, sph[c * NT], sx[c * NT]);             // x += alpha p^ now, + omega s^ below
                sph[c * NT] = zi;                                             // (p^ is done with: the column keeps s^)
                const double rh = srh[c * NT];
                d4[0] = fma(ti, r[c], d4[0]); d4[1] = fma(ti, ti, d4[1]);
                d4[2] = fma(rh, ti, d4[2]); d4[3] = fma(rh, r[c], d4[3]);
            });
            tiny_sum<4, NW>(d4, red, parity);
            const double omega = d4[1] > 0.0 ? d4[0] / d4[1] : 0.0;
            const double rho_new = d4[3] - omega * d4[2];
            double beta = 0.0;
            bool stagnated = false;
            if (omega != 0.0 && rho != 0.0) beta = (rho_new / rho) * (alpha / omega);
            else if (d4[1] > 0.0) stagnated = true;
            if (!(beta == beta) || isinf(beta)) { beta = 0.0; stagnated = true; }
            double dr[1] = {0.0};
#pragma unroll
            for (int c = 0; c < NS; ++c) {
                sx[c * NT] = fma(omega, sph[c * NT], sx[c * NT]);
                const double rn = fma(-omega, st[c * NT], r[c]);
                r[c] = rn;
                setp(c, fma(beta, fma(-omega, sv[c * NT], getp(c)), rn));
                dr[0] = fma(rn, rn, dr[0]);
            }
            zero_z();
            tiny_sum<1, NW>(dr, red, parity);            // publishes z = 0
            rr = dr[0];
            ++iters;
            rho = rho_new;
            if (!(rr == rr) || isinf(rr)) { flags |= FL_NAN; break; }
            if (rr <= M.tol2 * bb) { flags |= FL_CONVERGED; break; }
            if (stagnated) { breakdown = true; break; }
        }
        if ((flags & (FL_CONVERGED | FL_NAN)) || iters >= M.max_iter) break;
        if (breakdown && restarts < 3) { ++restarts; __syncthreads(); continue; }     // new shadow residual from the current iterate
        if (breakdown) flags |= FL_BREAKDOWN;
        break;
    }
#pragma unroll
    for (int c = 0; c < NS; ++c) if (CHIP_ON(c)) xg[(size_t)CHIP_ROW(c) * K] = sx[c * NT];
#undef CHIP_ON
#undef CHIP_ROW
    if (threadIdx.x == 0) {
        M.colflags[k] = flags; M.coliters[k] = iters;
        M.sc[SC_BNORM2 * K + k] = bb; M.sc[SC_RNORM2 * K + k] = rr;
        atomicMax(&stats->max_iterations, iters);
        atomicAdd(&stats->sum_iterations, (unsigned long long)iters);
        atomicOr(&stats->flags_or, flags & (FL_BREAKDOWN | FL_NAN));
        atomicMax(&stats->restarts, restarts);
        if (!(flags & FL_CONVERGED)) atomicAdd(&stats->not_converged, 1);
        const double rel2 = bb > 0.0 ? rr / bb : (rr > 0.0 ? __longlong_as_double(0x7ff0000000000000LL) : 0.0);
        if (rel2 == rel2) atomicMax(&stats->max_relres2_bits, (unsigned long long)__double_as_longlong(rel2));
    }
}
}  // namespace cwr
