// sm_100a kernels of the transport step.  fp64 throughout; HBM-bound gather/stream work, so the
// design rules are: coalescing (constituents interleaved, x[row*K + k]: one 128 B line per row at
// K = 16, moved as 128-bit double2 per lane), a fixed-width row-major ELL matrix (one int4 + two
// double2 broadcast loads per row, no rowptr dependency in front of the gathers), persistent grids
// sized from the SM count, fused dot products with a deterministic two-level reduction (no
// floating-point atomics anywhere), and per-step parameters read from a device-resident struct so
// that a step is the same launch sequence every time.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cwr {

constexpr int kThreads = 256;
constexpr int kMaxK = 128;      // constituents per handle
constexpr int kMaxDots = 4;
#ifndef CWR_SPMM_MIN_BLOCKS
#define CWR_SPMM_MIN_BLOCKS 4   // resident CTAs per SM the SpMM kernels are compiled for (register cap 64)
#endif

// Per-step pointers/values; written by k_set_step and read by every kernel of the step.
struct StepParams {
    const float* adv_t;      // (E)  advection_coeff[t]       device edge order
    const double* cdiff_t;   // (E)  coeff_to_diffusion[t]
    const float* vol_t;      // (n)  volume[t]   real cells, device cell order
    const float* vol_t1;     // (n)  volume[t+1]
    const float* adv_t1;     // (E)  advection_coeff[t+1]
    const double* cdiff_t1;  // (E)  coeff_to_diffusion[t+1]
    const float* velg_t1;    // (E_g) edge_velocity[t+1] of ghost edges
    const double* bc_t1;     // (G,K) input_array[t+1][ghost cells]
    const double* state_t;   // (n,K) c[t]
    double* state_t1;        // (n,K) c[t+1]  (the solver iterates in place on it)
    double dt;               // dt[t]
    int t;
    int apply_ic;            // t == 0: input_array[0] overrides c~ where non-zero (linalg.py:199-200)
};

enum ScalarRow { SC_RHO = 0, SC_ALPHA, SC_OMEGA, SC_BETA, SC_BNORM2, SC_RNORM2, SC_RHATV, SC_ROWS };
enum ColFlag { FL_CONVERGED = 1, FL_BREAKDOWN = 2, FL_NAN = 4, FL_ZERO_RHS = 8, FL_PENDING = 16 };

struct SolverCtl {
    int all_done;       // every column converged / failed / fixed
    int iter;           // BiCGSTAB iterations done in this solve
    int flags_or;       // OR of column flags that matter to the host (breakdown, nan)
    int singular;       // a zero diagonal was met during assembly
    int hit_max_iter;
    int pad[3];
    unsigned ticket[4]; // last-block tickets (one per kernel family)
};

struct DeviceModel {
    int n, K, E, E_int, E_g, G, nb, W;     // W = ELL width (multiple of 4)
    const int32_t* ell_col;   // (n,W) neighbour row, padded with the row itself
    const int32_t* ell_code;  // (n,W) (e' << 1) | side, -1 = padding
    const int32_t* f1p; const int32_t* f2p;
    const int32_t* bcell; const int32_t* bptr; const int32_t* bedge;
    double* val;        // (n,W) off-diagonals of D^-1 A
    double* diag;       // (n)   D
    double* gdiag;      // (n)   ghost-edge diagonal terms (boundary cells only, 0 elsewhere)
    const double* ic;   // (n,K) input_array[0][0:n]
    double *b, *r, *rhat, *p, *v, *tt, *ph, *sh, *tmp, *xc;   // (n,K) work vectors (b is the row-scaled RHS)
    double* partials;   // (grid, kMaxDots, K)
    double* sc;         // (SC_ROWS, K) per-column scalars
    int* colflags;      // (K)
    int* coliters;      // (K)
    SolverCtl* ctl;
    const StepParams* sp;
    double* flux;       // (3, E, K) advection / diffusion / total mass flux of the last step
    double* bsum;       // (3, E_g, K) running total / in / out sums on ghost edges
    double tol2;        // rtol^2
    double diffusion_coefficient;
    int max_iter;
    int want_flux;
};

__global__ void k_set_step(StepParams p, StepParams* dst, SolverCtl* ctl) {
    *dst = p;
    ctl->all_done = 0; ctl->iter = 0; ctl->flags_or = 0; ctl->hit_max_iter = 0;
}

// ---------------------------------------------------------------------------------------------
// upload helpers: reference order -> device order
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_gather(T* __restrict__ dst, const T* __restrict__ src, const int32_t* __restrict__ idx, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[idx[i]];
}

__global__ void k_scatter_column(double* __restrict__ dst, const double* __restrict__ src,
                                 const int32_t* __restrict__ old_of_new, int n, int K, int k) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        dst[(size_t)i * K + k] = src[old_of_new[i]];
}

// bc[(t*G + g)*K + k] = input[t*F + n + g]
__global__ void k_scatter_bc(double* __restrict__ bc, const double* __restrict__ input, int T, int F, int n, int G, int K, int k) {
    size_t total = (size_t)T * G;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t t = i / G, g = i % G;
        bc[i * K + k] = input[t * F + n + g];
    }
}

__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7ff8000000000000LL); }

// out[F]: real cells from the interleaved state (reference order), ghost cells: BC value or NaN
__global__ void k_extract_state(double* __restrict__ out, const double* __restrict__ state, const double* __restrict__ bc_t,
                                const int32_t* __restrict__ new_of_old, int n, int F, int K, int k) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < F; j += gridDim.x * blockDim.x) {
        double v;
        if (j < n) v = state[(size_t)new_of_old[j] * K + k];
        else {
            v = bc_t ? bc_t[(size_t)(j - n) * K + k] : 0.0;
            if (v == 0.0) v = qnan();   // transport.py:258-264: unset ghost cells stay NaN
        }
        out[j] = v;
    }
}

// out[(k, j)] for all constituents, real cells only, reference order
__global__ void k_extract_all(double* __restrict__ out, const double* __restrict__ state,
                              const int32_t* __restrict__ new_of_old, int n, int K) {
    size_t total = (size_t)n * K;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int k = (int)(i / n), j = (int)(i % n);
        out[i] = state[(size_t)new_of_old[j] * K + k];
    }
}

__global__ void k_scatter_all(double* __restrict__ state, const double* __restrict__ src, const uint8_t* __restrict__ mask,
                              const int32_t* __restrict__ new_of_old, int n, int K) {
    size_t total = (size_t)n * K;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int k = (int)(i / n), j = (int)(i % n);
        if (mask == nullptr || mask[k]) state[(size_t)new_of_old[j] * K + k] = src[i];
    }
}

// out[e] (reference edge order) = flux[which][e'][k]
__global__ void k_extract_flux(double* __restrict__ out, const double* __restrict__ flux, const int32_t* __restrict__ einv,
                               int E, int K, int k) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x)
        out[e] = flux[(size_t)einv[e] * K + k];
}

// ---------------------------------------------------------------------------------------------
// "next" row N1: derive adv / cdiff from the raw HEC-RAS arrays (reference utilities.py:513-541)
//   adv = face_flow * sign(|vel|)  (f32);  area = adv / vel, NaN -> 0 (f32);
//   cdiff = f64(f32(area * D)) / dist
// input in reference edge order, output in device edge order.
// ---------------------------------------------------------------------------------------------
__global__ void k_derive(float* __restrict__ adv, double* __restrict__ cdiff, float* __restrict__ velg,
                         const float* __restrict__ flow, const float* __restrict__ vel, const double* __restrict__ dist,
                         const int32_t* __restrict__ eperm, int E, int E_int, float Df) {
    for (int ep = blockIdx.x * blockDim.x + threadIdx.x; ep < E; ep += gridDim.x * blockDim.x) {
        int e = eperm[ep];
        float q = flow[e], u = vel[e];
        float au = fabsf(u);
        float sg = (au > 0.f) ? 1.f : ((au == 0.f) ? 0.f : au);     // sign(|u|); NaN stays NaN
        float a = __fmul_rn(q, sg);
        float area = __fdiv_rn(a, u);
        if (area != area) area = 0.f;                                // fillna(0)
        float ad = __fmul_rn(area, Df);
        adv[ep] = a;
        cdiff[ep] = (double)ad / dist[e];
        if (ep >= E_int) velg[ep - E_int] = u;
    }
}

__global__ void k_dist(double* __restrict__ dist, const double* __restrict__ fx, const double* __restrict__ fy,
                       const int32_t* __restrict__ f1, const int32_t* __restrict__ f2, int E) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
        double dx = fx[f1[e]] - fx[f2[e]], dy = fy[f1[e]] - fy[f2[e]];
        dist[e] = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));   // no fma contraction: matches numpy
    }
}

// ---------------------------------------------------------------------------------------------
// LHS assembly  (reference linalg.py:34-156 + transport.py:215-218)
// ---------------------------------------------------------------------------------------------
// ghost-edge diagonal terms: sum over a boundary cell's ghost edges of cdiff + max(adv, 0)
// (linalg.py:92-97 and 113-115 applied to ghost edges).
__global__ void k_boundary_diag(DeviceModel M) {
    const StepParams& sp = *M.sp;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < M.nb; b += gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int j = M.bptr[b]; j < M.bptr[b + 1]; ++j) {
            int e = M.bedge[j];
            s += sp.cdiff_t[e] + fmax((double)sp.adv_t[e], 0.0);
        }
        M.gdiag[M.bcell[b]] = s;
    }
}

// One thread per row: diagonal D_i and the row's off-diagonals, written once into their fixed
// slots, already divided by D_i (row-scaled system D^-1 A, unit diagonal implied).
//   internal edge e = (P = f1, N = f2), a = adv[t,e], d = cdiff[t,e]:
//     A[P,N] = -d + min(a,0)   A[N,P] = -d - max(a,0)
//     A[P,P] += d + max(a,0)   A[N,N] += d - min(a,0)
//   A[i,i] += vol[t+1,i]/dt[t]  (+1 if vol[t+1,i] == 0, linalg.py:66,77-81)
__global__ void __launch_bounds__(kThreads) k_assemble(DeviceModel M) {
    const StepParams& sp = *M.sp;
    const double dt = sp.dt;
    const int W = M.W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M.n; i += gridDim.x * blockDim.x) {
        const float vol = sp.vol_t1[i];
        double diag = (vol == 0.f ? 1.0 : 0.0) + (double)vol / dt + M.gdiag[i];
        const int32_t* code = M.ell_code + (size_t)i * W;
        double* val = M.val + (size_t)i * W;
        for (int w = 0; w < W; w += 4) {
            const int4 c4 = *reinterpret_cast<const int4*>(code + w);
            const int cs[4] = {c4.x, c4.y, c4.z, c4.w};
            double off[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                off[u] = 0.0;
                if (cs[u] >= 0) {
                    const double a = (double)sp.adv_t[cs[u] >> 1];
                    const double d = sp.cdiff_t[cs[u] >> 1];
                    if (cs[u] & 1) { off[u] = -d - fmax(a, 0.0); diag += d - fmin(a, 0.0); }
                    else           { off[u] = -d + fmin(a, 0.0); diag += d + fmax(a, 0.0); }
                }
            }
            *reinterpret_cast<double2*>(val + w) = make_double2(off[0], off[1]);
            *reinterpret_cast<double2*>(val + w + 2) = make_double2(off[2], off[3]);
        }
        const double inv = 1.0 / diag;
        for (int w = 0; w < W; w += 2) {
            double2 o = *reinterpret_cast<double2*>(val + w);
            o.x *= inv; o.y *= inv;
            *reinterpret_cast<double2*>(val + w) = o;
        }
        M.diag[i] = diag;
        if (diag == 0.0) M.ctl->singular = 1;
    }
}

// ---------------------------------------------------------------------------------------------
// (row, column) thread mapping shared by all vector kernels: KC lanes per row, each lane owns VEC
// adjacent columns (VEC = 2: 128-bit accesses); rows are dealt to lane groups grid-stride.
// ---------------------------------------------------------------------------------------------
template <int VEC> struct Vd { double a[VEC]; };

template <int VEC>
__device__ __forceinline__ Vd<VEC> ldv(const double* p) {
    Vd<VEC> r;
    if (VEC == 2) { const double2 t = *reinterpret_cast<const double2*>(p); r.a[0] = t.x; r.a[VEC - 1] = t.y; }
    else r.a[0] = *p;
    return r;
}
template <int VEC>
__device__ __forceinline__ void stv(double* p, const Vd<VEC>& v) {
    if (VEC == 2) *reinterpret_cast<double2*>(p) = make_double2(v.a[0], v.a[VEC - 1]);
    else *p = v.a[0];
}

// RHS  (reference linalg.py:177-275), K constituents at once, row-scaled by 1/D; warm start x0 = c~
template <int KC, int VEC>
__global__ void __launch_bounds__(kThreads) k_rhs(DeviceModel M) {
    const StepParams& sp = *M.sp;
    const int K = M.K, lane = threadIdx.x % KC, group = threadIdx.x / KC, GPB = kThreads / KC;
    const double dt = sp.dt;
    for (int c = lane * VEC; c < K; c += KC * VEC)
        for (int i = blockIdx.x * GPB + group; i < M.n; i += gridDim.x * GPB) {
            const size_t idx = (size_t)i * K + c;
            Vd<VEC> conc = ldv<VEC>(sp.state_t + idx);
            if (sp.apply_ic) {
                const Vd<VEC> ic = ldv<VEC>(M.ic + idx);
#pragma unroll
                for (int q = 0; q < VEC; ++q) if (ic.a[q] != 0.0) conc.a[q] = ic.a[q];
            }
            const double vol = (double)sp.vol_t[i], dg = M.diag[i];
            Vd<VEC> bb;
#pragma unroll
            for (int q = 0; q < VEC; ++q) bb.a[q] = (vol * conc.a[q] / dt) / dg;      // linalg.py:239
            stv<VEC>(M.b + idx, bb);
            stv<VEC>(sp.state_t1 + idx, conc);
        }
}

// Boundary cells: b_i = load + ghost_in + ghost_out with the reference's selection and
// last-edge-wins assignment (linalg.py:349-351, 372-378, 390).  One lane per (cell, column).
__global__ void __launch_bounds__(kThreads) k_boundary_rhs(DeviceModel M) {
    const StepParams& sp = *M.sp;
    const int K = M.K;
    const double dt = sp.dt;
    const bool has_diffusion = M.diffusion_coefficient != 0.0;
    const size_t total = (size_t)M.nb * K;
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(q / K), c = (int)(q % K);
        const int i = M.bcell[b];
        double m_in = 0.0, ca_in = 0.0, cd_in = 0.0, m_out = 0.0, cd_out = 0.0;
        for (int j = M.bptr[b]; j < M.bptr[b + 1]; ++j) {
            const int e = M.bedge[j];
            const float u = sp.velg_t1[e - M.E_int];
            const double bc = sp.bc_t1[(size_t)(M.f2p[e] - M.n) * K + c];
            const double cd = has_diffusion ? fabs(sp.cdiff_t1[e]) : 0.0;
            if (u < 0.f) { m_in = bc; ca_in = fabs((double)sp.adv_t1[e]); cd_in = cd; }
            if (u > 0.f) { m_out = bc; cd_out = cd; }
        }
        const size_t idx = (size_t)i * K + c;
        const double conc = sp.state_t1[idx];                    // c~ written by k_rhs
        const double load = (double)sp.vol_t[i] * conc / dt;
        const double rhs = load + (ca_in + cd_in) * m_in + cd_out * m_out;
        M.b[idx] = rhs / M.diag[i];
    }
}

// ---------------------------------------------------------------------------------------------
// deterministic block / grid reduction of per-column dot products
// ---------------------------------------------------------------------------------------------
// acc[d*VEC + q]: dot d of column (chunk*KC + lane)*VEC + q.  block_out: [kMaxDots][K] of this block.
template <int ND, int KC, int VEC>
__device__ __forceinline__ void block_dots(double (&acc)[ND * VEC], double* smem, double* block_out, int K, int chunk) {
    constexpr int NA = ND * VEC;
#pragma unroll
    for (int off = KC; off < 32; off <<= 1)
#pragma unroll
        for (int d = 0; d < NA; ++d) acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], off);
    const int wl = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = kThreads / 32;
    if (wl < KC)
#pragma unroll
        for (int d = 0; d < NA; ++d) smem[(warp * NA + d) * KC + wl] = acc[d];
    __syncthreads();
    if (threadIdx.x < NA * KC) {
        const int a = threadIdx.x / KC, l = threadIdx.x % KC;
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += smem[(w * NA + a) * KC + l];
        const int d = a / VEC, q = a % VEC;
        const int c = (chunk * KC + l) * VEC + q;
        if (c < K) block_out[d * K + c] = s;
    }
    __syncthreads();
}

// true in exactly one block: the last one to arrive.  All of the grid's partials are visible to it.
__device__ __forceinline__ bool last_block_arrives(unsigned* ticket) {
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
        if (is_last) *ticket = 0;
    }
    __syncthreads();
    if (is_last) __threadfence();
    return is_last != 0;
}

// sum the grid's partials in block order -> tot[d*K + k] (shared memory)
// All 256 threads take part: each (dot, column) pair is summed by a team of threads over interleaved
// block ranges, then the team's partial sums are combined in a fixed order -> still deterministic.
template <int ND>
__device__ __forceinline__ void grid_totals(const double* partials, double* tot, int K) {
    __shared__ double team_sum[kThreads];
    const int pairs = ND * K;
    int team = 1;
    while (team * 2 * pairs <= kThreads && team < 64) team *= 2;     // threads per pair (power of two)
    for (int base = 0; base < pairs; base += kThreads / team) {
        const int pair = base + threadIdx.x / team, member = threadIdx.x % team;
        double s = 0.0;
        if (pair < pairs) {
            const int d = pair / K, k = pair % K;
            for (unsigned b = member; b < gridDim.x; b += team) s += __ldcg(&partials[((size_t)b * kMaxDots + d) * K + k]);
        }
        team_sum[threadIdx.x] = s;
        __syncthreads();
        if (pair < pairs && member == 0) {
            double t = 0.0;
            for (int m = 0; m < team; ++m) t += team_sum[threadIdx.x + m];
            tot[pair] = t;
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void publish_done(DeviceModel& M, int K) {
    int done = 1, flags = 0;
    for (int k = 0; k < K; ++k) {
        const int f = M.colflags[k];
        flags |= f;
        if (!(f & (FL_CONVERGED | FL_BREAKDOWN | FL_NAN)) || (f & FL_PENDING)) done = 0;
    }
    M.ctl->flags_or = flags;
    if (M.ctl->iter >= M.max_iter && !done) { M.ctl->hit_max_iter = 1; done = 1; }
    M.ctl->all_done = done;
}

// ---------------------------------------------------------------------------------------------
// SpMM over the row-scaled matrix  A = I + L  (L = off-diagonals in ELL; N := -L is the Jacobi
// iteration matrix), fused with the BiCGSTAB step that consumes it.  z is gathered, u is the
// row's own vector.
//   INIT : r = b - (x + L x)          ; rhat = p = r ; dots (r,r), (b,b)
//   JAC  : out = u - L z              (one step of the m-step Jacobi preconditioner: out = u + N z)
//   AV   : v = z + L z                ; dot (rhat, v)                      -> alpha
//   AT   : t = z + L z                ; dots (t,s),(t,t),(rhat,t),(rhat,s) -> omega, rho', beta
//   PLAIN: y = z + L z                (timing / tests)
// ---------------------------------------------------------------------------------------------
enum SpmmMode { MODE_INIT = 0, MODE_AV = 1, MODE_AT = 2, MODE_JAC = 3, MODE_PLAIN = 4 };

template <int KC, int VEC, int MODE>
__global__ void __launch_bounds__(kThreads, CWR_SPMM_MIN_BLOCKS) k_spmm(DeviceModel M, const double* __restrict__ zin,
                                                   const double* __restrict__ uin, double* __restrict__ out) {
    constexpr int ND = MODE == MODE_INIT ? 2 : MODE == MODE_AV ? 1 : MODE == MODE_AT ? 4 : 1;
    constexpr bool HAS_DOTS = MODE == MODE_INIT || MODE == MODE_AV || MODE == MODE_AT;
    __shared__ double smem[HAS_DOTS ? (kThreads / 32) * kMaxDots * 2 * 32 : 1];
    __shared__ double tot[HAS_DOTS ? kMaxDots * kMaxK : 1];
    if (MODE == MODE_AV || MODE == MODE_AT || MODE == MODE_JAC) { if (M.ctl->all_done) return; }
    const int K = M.K, n = M.n, W = M.W;
    const int lane = threadIdx.x % KC, group = threadIdx.x / KC, GPB = kThreads / KC;
    const int32_t* __restrict__ ecol = M.ell_col;
    const double* __restrict__ eval = M.val;
    if (MODE == MODE_INIT) zin = M.sp->state_t1;
    const int nchunk = (K + KC * VEC - 1) / (KC * VEC);
    for (int chunk = 0; chunk < nchunk; ++chunk) {
        const int c = (chunk * KC + lane) * VEC;
        const bool active = c < K;
        double acc[ND * VEC];
#pragma unroll
        for (int d = 0; d < ND * VEC; ++d) acc[d] = 0.0;
        if (active) {
            // Software pipeline: the next row's column indices and values (the loads the gathers depend
            // on) are fetched while this row's gathers are in flight, so a row costs one memory latency,
            // not two.
            const int stride = gridDim.x * GPB;
            int i = blockIdx.x * GPB + group;
            int4 c4n = make_int4(0, 0, 0, 0);
            double2 v01n = make_double2(0.0, 0.0), v23n = v01n;
            if (i < n) {
                c4n = *reinterpret_cast<const int4*>(ecol + (size_t)i * W);
                v01n = *reinterpret_cast<const double2*>(eval + (size_t)i * W);
                v23n = *reinterpret_cast<const double2*>(eval + (size_t)i * W + 2);
            }
            for (; i < n; i += stride) {
                const size_t idx = (size_t)i * K + c;
                const int4 c4 = c4n;
                const double2 v01 = v01n, v23 = v23n;
                const Vd<VEC> x0 = ldv<VEC>(zin + (size_t)c4.x * K + c);
                const Vd<VEC> x1 = ldv<VEC>(zin + (size_t)c4.y * K + c);
                const Vd<VEC> x2 = ldv<VEC>(zin + (size_t)c4.z * K + c);
                const Vd<VEC> x3 = ldv<VEC>(zin + (size_t)c4.w * K + c);
                // the row's own operands are issued now as well, so nothing waits behind the gathers
                const Vd<VEC> own = ldv<VEC>((MODE == MODE_JAC ? uin : zin) + idx);
                Vd<VEC> aux1 = own, aux2 = own;
                if (MODE == MODE_INIT) aux1 = ldv<VEC>(M.b + idx);
                if (MODE == MODE_AV || MODE == MODE_AT) aux1 = ldv<VEC>(M.rhat + idx);
                if (MODE == MODE_AT) aux2 = ldv<VEC>(M.r + idx);          // s lives in the r buffer
                const int inext = i + stride;
                if (inext < n) {
                    c4n = *reinterpret_cast<const int4*>(ecol + (size_t)inext * W);
                    v01n = *reinterpret_cast<const double2*>(eval + (size_t)inext * W);
                    v23n = *reinterpret_cast<const double2*>(eval + (size_t)inext * W + 2);
                }
                Vd<VEC> s;
#pragma unroll
                for (int q = 0; q < VEC; ++q)
                    s.a[q] = fma(v23.y, x3.a[q], fma(v23.x, x2.a[q], fma(v01.y, x1.a[q], v01.x * x0.a[q])));
                for (int w = 4; w < W; w += 4) {      // rows wider than 4 (not pipelined)
                    const int4 d4 = *reinterpret_cast<const int4*>(ecol + (size_t)i * W + w);
                    const double2 w01 = *reinterpret_cast<const double2*>(eval + (size_t)i * W + w);
                    const double2 w23 = *reinterpret_cast<const double2*>(eval + (size_t)i * W + w + 2);
                    const Vd<VEC> y0 = ldv<VEC>(zin + (size_t)d4.x * K + c);
                    const Vd<VEC> y1 = ldv<VEC>(zin + (size_t)d4.y * K + c);
                    const Vd<VEC> y2 = ldv<VEC>(zin + (size_t)d4.z * K + c);
                    const Vd<VEC> y3 = ldv<VEC>(zin + (size_t)d4.w * K + c);
#pragma unroll
                    for (int q = 0; q < VEC; ++q)
                        s.a[q] = fma(w23.y, y3.a[q], fma(w23.x, y2.a[q], fma(w01.y, y1.a[q], fma(w01.x, y0.a[q], s.a[q]))));
                }
                if (MODE == MODE_JAC) {
                    Vd<VEC> o;
#pragma unroll
                    for (int q = 0; q < VEC; ++q) o.a[q] = own.a[q] - s.a[q];
                    stv<VEC>(out + idx, o);
                } else {
                    Vd<VEC> y;
#pragma unroll
                    for (int q = 0; q < VEC; ++q) y.a[q] = own.a[q] + s.a[q];
                    if (MODE == MODE_INIT) {
                        const Vd<VEC> bi = aux1;
                        Vd<VEC> r;
#pragma unroll
                        for (int q = 0; q < VEC; ++q) {
                            r.a[q] = bi.a[q] - y.a[q];
                            acc[0 * VEC + q] = fma(r.a[q], r.a[q], acc[0 * VEC + q]);
                            acc[1 * VEC + q] = fma(bi.a[q], bi.a[q], acc[1 * VEC + q]);
                        }
                        stv<VEC>(M.r + idx, r); stv<VEC>(M.rhat + idx, r); stv<VEC>(M.p + idx, r);
                    } else if (MODE == MODE_AV) {
                        const Vd<VEC> rh = aux1;
                        stv<VEC>(M.v + idx, y);
#pragma unroll
                        for (int q = 0; q < VEC; ++q) acc[q] = fma(rh.a[q], y.a[q], acc[q]);
                    } else if (MODE == MODE_AT) {
                        const Vd<VEC> rh = aux1, sv = aux2;
                        stv<VEC>(M.tt + idx, y);
#pragma unroll
                        for (int q = 0; q < VEC; ++q) {
                            acc[0 * VEC + q] = fma(y.a[q], sv.a[q], acc[0 * VEC + q]);
                            acc[1 * VEC + q] = fma(y.a[q], y.a[q], acc[1 * VEC + q]);
                            acc[2 * VEC + q] = fma(rh.a[q], y.a[q], acc[2 * VEC + q]);
                            acc[3 * VEC + q] = fma(rh.a[q], sv.a[q], acc[3 * VEC + q]);
                        }
                    } else {
                        stv<VEC>(out + idx, y);
                    }
                }
            }
        }
        if (HAS_DOTS) block_dots<ND, KC, VEC>(acc, smem, M.partials + (size_t)blockIdx.x * kMaxDots * K, K, chunk);
    }
    if (!HAS_DOTS) return;
    if (!last_block_arrives(&M.ctl->ticket[MODE])) return;
    grid_totals<ND>(M.partials, tot, K);
    double* sc = M.sc;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        int f = M.colflags[k];
        if (MODE == MODE_INIT) {
            const double rr = tot[0 * K + k], bb = tot[1 * K + k];
            sc[SC_RHO * K + k] = rr; sc[SC_BNORM2 * K + k] = bb; sc[SC_RNORM2 * K + k] = rr;
            sc[SC_ALPHA * K + k] = 0.0; sc[SC_OMEGA * K + k] = 0.0; sc[SC_BETA * K + k] = 0.0;
            f &= ~(FL_CONVERGED | FL_BREAKDOWN | FL_PENDING | FL_ZERO_RHS | FL_NAN);
            if (!(rr == rr) || !(bb == bb) || isinf(rr) || isinf(bb)) f |= FL_NAN | FL_PENDING;
            else if (bb == 0.0 && rr != 0.0) f |= FL_ZERO_RHS | FL_PENDING;     // b == 0  =>  x = 0
            else if (rr <= M.tol2 * bb) { f |= FL_CONVERGED; M.coliters[k] = M.ctl->iter; }
        } else if (MODE == MODE_AV) {
            const double rv = tot[k];
            sc[SC_RHATV * K + k] = rv;
            double alpha = 0.0;
            if (!(f & (FL_CONVERGED | FL_BREAKDOWN | FL_NAN | FL_PENDING))) {
                if (rv == 0.0 || !(rv == rv)) f |= FL_BREAKDOWN;
                else alpha = sc[SC_RHO * K + k] / rv;
            }
            sc[SC_ALPHA * K + k] = alpha;
        } else {
            const double ts = tot[0 * K + k], t2 = tot[1 * K + k], rt = tot[2 * K + k], rs = tot[3 * K + k];
            double omega = 0.0, beta = 0.0;
            if (!(f & (FL_CONVERGED | FL_BREAKDOWN | FL_NAN | FL_PENDING))) {
                const double rho = sc[SC_RHO * K + k], alpha = sc[SC_ALPHA * K + k];
                omega = t2 > 0.0 ? ts / t2 : 0.0;
                const double rho_new = rs - omega * rt;          // (rhat, s - omega t)
                if (omega != 0.0 && rho != 0.0) beta = (rho_new / rho) * (alpha / omega);
                else if (t2 > 0.0) f |= FL_BREAKDOWN;            // omega == 0 with s != 0: stagnation
                if (!(beta == beta) || isinf(beta)) { beta = 0.0; f |= FL_BREAKDOWN; }
                sc[SC_RHO * K + k] = rho_new;
            }
            sc[SC_OMEGA * K + k] = omega; sc[SC_BETA * K + k] = beta;
        }
        M.colflags[k] = f;
    }
    __syncthreads();
    if (MODE == MODE_INIT && threadIdx.x == 0) publish_done(M, K);
}

// s = r - alpha v  (in place on r)
template <int KC, int VEC>
__global__ void __launch_bounds__(kThreads) k_update_s(DeviceModel M) {
    if (M.ctl->all_done) return;
    const int K = M.K, lane = threadIdx.x % KC, group = threadIdx.x / KC, GPB = kThreads / KC;
    for (int c = lane * VEC; c < K; c += KC * VEC) {
        double alpha[VEC];
        bool any = false;
#pragma unroll
        for (int q = 0; q < VEC; ++q) { alpha[q] = M.sc[SC_ALPHA * K + c + q]; any |= alpha[q] != 0.0; }
        if (!any) continue;        // frozen columns: s = r
        for (int i = blockIdx.x * GPB + group; i < M.n; i += gridDim.x * GPB) {
            const size_t idx = (size_t)i * K + c;
            Vd<VEC> r = ldv<VEC>(M.r + idx);
            const Vd<VEC> v = ldv<VEC>(M.v + idx);
#pragma unroll
            for (int q = 0; q < VEC; ++q) r.a[q] = fma(-alpha[q], v.a[q], r.a[q]);
            stv<VEC>(M.r + idx, r);
        }
    }
}

// x += alpha ph + omega sh ; r = s - omega t ; p = r + beta (p - omega v) ; dot (r,r); convergence
// (ph, sh: preconditioned p and s; equal to p and s when the preconditioner is the identity)
template <int KC, int VEC>
__global__ void __launch_bounds__(kThreads) k_update_xrp(DeviceModel M, const double* __restrict__ ph,
                                                         const double* __restrict__ sh) {
    __shared__ double smem[(kThreads / 32) * kMaxDots * 2 * 32];
    __shared__ double tot[kMaxDots * kMaxK];
    if (M.ctl->all_done) return;
    const int K = M.K, lane = threadIdx.x % KC, group = threadIdx.x / KC, GPB = kThreads / KC;
    double* __restrict__ x = M.sp->state_t1;
    const int nchunk = (K + KC * VEC - 1) / (KC * VEC);
    for (int chunk = 0; chunk < nchunk; ++chunk) {
        const int c = (chunk * KC + lane) * VEC;
        double acc[VEC];
#pragma unroll
        for (int q = 0; q < VEC; ++q) acc[q] = 0.0;
        if (c < K) {
            int f[VEC]; double alpha[VEC], omega[VEC], beta[VEC];
            bool any_pending = false, any_active = false;
#pragma unroll
            for (int q = 0; q < VEC; ++q) {
                f[q] = M.colflags[c + q];
                alpha[q] = M.sc[SC_ALPHA * K + c + q]; omega[q] = M.sc[SC_OMEGA * K + c + q]; beta[q] = M.sc[SC_BETA * K + c + q];
                any_pending |= (f[q] & FL_PENDING) != 0;
                any_active |= !(f[q] & (FL_CONVERGED | FL_BREAKDOWN | FL_NAN | FL_PENDING));
            }
            if (any_pending || any_active)
                for (int i = blockIdx.x * GPB + group; i < M.n; i += gridDim.x * GPB) {
                    const size_t idx = (size_t)i * K + c;
                    Vd<VEC> xv = ldv<VEC>(x + idx), rv = ldv<VEC>(M.r + idx), pv = ldv<VEC>(M.p + idx);
                    const Vd<VEC> tv = ldv<VEC>(M.tt + idx), vv = ldv<VEC>(M.v + idx);
                    const Vd<VEC> phv = ldv<VEC>(ph + idx), shv = ldv<VEC>(sh + idx);
#pragma unroll
                    for (int q = 0; q < VEC; ++q) {
                        if (f[q] & FL_PENDING) {
                            xv.a[q] = (f[q] & FL_NAN) ? qnan() : 0.0; rv.a[q] = 0.0; pv.a[q] = 0.0;
                        } else if (!(f[q] & (FL_CONVERGED | FL_BREAKDOWN | FL_NAN))) {
                            xv.a[q] = fma(omega[q], shv.a[q], fma(alpha[q], phv.a[q], xv.a[q]));
                            const double rn = fma(-omega[q], tv.a[q], rv.a[q]);
                            rv.a[q] = rn;
                            pv.a[q] = fma(beta[q], fma(-omega[q], vv.a[q], pv.a[q]), rn);
                            acc[q] = fma(rn, rn, acc[q]);
                        }
                    }
                    stv<VEC>(x + idx, xv); stv<VEC>(M.r + idx, rv); stv<VEC>(M.p + idx, pv);
                }
        }
        block_dots<1, KC, VEC>(acc, smem, M.partials + (size_t)blockIdx.x * kMaxDots * K, K, chunk);
    }
    if (!last_block_arrives(&M.ctl->ticket[3])) return;
    grid_totals<1>(M.partials, tot, K);
    if (threadIdx.x == 0) M.ctl->iter += 1;
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        int f = M.colflags[k];
        if (f & FL_PENDING) {
            f &= ~FL_PENDING;
            if (f & FL_ZERO_RHS) { f |= FL_CONVERGED; M.sc[SC_RNORM2 * K + k] = 0.0; M.coliters[k] = M.ctl->iter; }
        } else if (!(f & (FL_CONVERGED | FL_BREAKDOWN | FL_NAN))) {
            const double rr = tot[k];
            M.sc[SC_RNORM2 * K + k] = rr;
            if (!(rr == rr) || isinf(rr)) f |= FL_NAN;
            else if (rr <= M.tol2 * M.sc[SC_BNORM2 * K + k]) { f |= FL_CONVERGED; M.coliters[k] = M.ctl->iter; }
        }
        M.colflags[k] = f;
    }
    __syncthreads();
    if (threadIdx.x == 0) publish_done(M, K);
}

// ---------------------------------------------------------------------------------------------
// mass flux across every edge  (reference transport.py:406-429) + running boundary sums
// (postproc_util.py:100-143).  c[t+1] of a ghost cell = its BC value, NaN when unset.
// ---------------------------------------------------------------------------------------------
template <int KC, int VEC>
__global__ void __launch_bounds__(kThreads) k_mass_flux(DeviceModel M) {
    const StepParams& sp = *M.sp;
    const int K = M.K, lane = threadIdx.x % KC, group = threadIdx.x / KC, GPB = kThreads / KC;
    const double dt = sp.dt;
    const double* __restrict__ x = sp.state_t1;
    const size_t EK = (size_t)M.E * K, GK = (size_t)M.E_g * K;
    for (int c = lane * VEC; c < K; c += KC * VEC)
        for (int e = blockIdx.x * GPB + group; e < M.E; e += gridDim.x * GPB) {
            const int P = M.f1p[e], N = M.f2p[e];
            const double a = (double)sp.adv_t[e], d = sp.cdiff_t[e];
            const Vd<VEC> cP = ldv<VEC>(x + (size_t)P * K + c);
            Vd<VEC> cN;
            if (N < M.n) cN = ldv<VEC>(x + (size_t)N * K + c);
            else {
                cN = ldv<VEC>(sp.bc_t1 + (size_t)(N - M.n) * K + c);
#pragma unroll
                for (int q = 0; q < VEC; ++q) if (cN.a[q] == 0.0) cN.a[q] = qnan();
            }
            Vd<VEC> fa, fd, ft;
#pragma unroll
            for (int q = 0; q < VEC; ++q) {
                fa.a[q] = __dmul_rn((a < 0.0 ? __dmul_rn(a, cN.a[q]) : __dmul_rn(a, cP.a[q])), dt);
                fd.a[q] = __dmul_rn(__dmul_rn(d, cN.a[q] - cP.a[q]), dt);
                ft.a[q] = fa.a[q] + fd.a[q];
            }
            const size_t o = (size_t)e * K + c;
            stv<VEC>(M.flux + o, fa); stv<VEC>(M.flux + EK + o, fd); stv<VEC>(M.flux + 2 * EK + o, ft);
            if (e >= M.E_int) {
                const size_t g = (size_t)(e - M.E_int) * K + c;
#pragma unroll
                for (int q = 0; q < VEC; ++q) {
                    const double f = ft.a[q];
                    M.bsum[g + q] += f;
                    M.bsum[GK + g + q] += (f <= 0.0) ? f : f * 0.0;
                    M.bsum[2 * GK + g + q] += (f >= 0.0) ? f : f * 0.0;
                }
            }
        }
}

// sum_i vol[i] * c[i,k]  -> out[k]   (single block, deterministic; postproc_util.py:36-59)
__global__ void k_mass_total(const float* __restrict__ vol, const double* __restrict__ state, int n, int K, int k,
                             double* __restrict__ out /* [2]: volume, mass */) {
    __shared__ double sv[kThreads], sm[kThreads];
    double v = 0.0, m = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double vi = (double)vol[i];
        v += vi; m = fma(vi, state[(size_t)i * K + k], m);
    }
    sv[threadIdx.x] = v; sm[threadIdx.x] = m;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) { sv[threadIdx.x] += sv[threadIdx.x + s]; sm[threadIdx.x] += sm[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = sv[0]; out[1] = sm[0]; }
}

}  // namespace cwr
