// np_tail.cuh -- K11: the coarse tail of the V-cycle in ONE kernel launch, one CTA, shared memory only.
//
// MLMG::mgVcycle visits every multigrid level; on the levels below ~17^3 nodes a kernel does a few
// hundred node updates and costs nothing but its launch / drain latency: 21 launches per level and V-cycle
// (16 sweeps, residual, restriction, interpolation, memset, copy) at 4-6 us each, ~0.34 ms of a 3.4 ms
// V-cycle at 256^3 for 0.03 % of the nodes (profiles/r1_phase_profile_256_rt.txt).  This kernel runs the
// whole down-leg, the bottom solve and the up-leg of those levels with __syncthreads() in place of kernel
// boundaries; every array of every tail level lives in shared memory for the duration.
//
// Semantics are those of the per-level kernels, restated node by node (this is the oracle's formulation):
//   * smoother (A.4): Gauss-Seidel inside 64 x 16 x tz tile-chunks, planes ascending, colours
//     c = (i&1) + 2(j&1) in order 0..3 inside a plane, previous-sweep values outside the tile-chunk
//     (np_smooth.cuh K4); tz is the level's chunk height from build_levels();
//   * residual (A.3), full-weighting restriction (A.5), sigma-weighted interpolation (A.6, `Interp`),
//     BiCGStab / CG bottom solve (A.10, bottom_bicgstab_body).
// Works for variable and constant sigma, isotropic or not, any BC combination (index maps of np_level.h).
#pragma once
#include "np_kernels.cuh"

namespace b200np_dev {

constexpr int TAIL_MAX_LEV = 8;
constexpr int TAIL_THREADS = 512;
constexpr int TAIL_MAX_NODES = 6144;        // per level: a colour step is one or two passes of the CTA
constexpr int TAIL_SMEM_DOUBLES = 27 * 1024;  // 216 KB of the 227 KB a CTA may own

struct TailLev {
    Lev g;                    // compact descriptor: px = nn[0], ps = nn[0] nn[1], cpx = n[0], cps = n[0] n[1], not distributed
    int a, b, r, s;           // offsets (doubles) of ping / pong / rhs / sigma in shared memory
    int tz;                   // smoother z-chunk of the level
    int cpx_g;                // pitch of the level's sigma array in global memory
    long long cps_g;
    const double* sigma_g;    // coarsened sigma (global), nullptr for constant sigma
};
struct TailPlan {
    int nlev;                 // tail levels; the last one is the multigrid bottom level
    int work;                 // offset of the bottom solver's 7 work vectors
    int px_io;                // pitch / plane stride of res_in and cor_out (global arrays of the first tail level)
    long long ps_io;
    const double* res_in;
    double* cor_out;
    int nu1, nu2, nsw;        // smooth calls before / after, sweeps per call
    int maxiter, singular, bottom_solver;
    double rtol, atol;
    int* info;
    TailLev lv[TAIL_MAX_LEV];
};

// 27 values around node (i,j,k): from `in` when the neighbour lies in the node's own tile-chunk, else from `out`
// (a neighbour reached through a periodic wrap or a reflection is outside, like a halo entry of the tile kernels)
__device__ __forceinline__ void tail_gather(const Lev& g, int tz, const double* __restrict__ in, const double* __restrict__ out,
                                            int i, int j, int k, double (&P)[3][3][3])
{
    int ix[3], jy[3], kz[3];
    bool bx[3], by[3], bz[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int ri = i - 1 + a, rj = j - 1 + a, rk = k - 1 + a;
        ix[a] = nmap(ri, g.n[0], g.per[0]); jy[a] = nmap(rj, g.n[1], g.per[1]); kz[a] = nmap(rk, g.n[2], g.per[2]);
        bx[a] = ri >= 0 && ri < g.nn[0] && ri / NP_TX == i / NP_TX;
        by[a] = rj >= 0 && rj < g.nn[1] && rj / NP_TY == j / NP_TY;
        bz[a] = rk >= 0 && rk < g.nn[2] && rk / tz == k / tz;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const int id = (kz[c] * g.nn[1] + jy[b]) * g.nn[0] + ix[a];
                P[c][b][a] = (bx[a] && by[b] && bz[c]) ? in[id] : out[id];
            }
}

template <bool VAR>
__device__ __forceinline__ double tail_Lphi(const Lev& g, const double (&P)[3][3][3], int i, int j, int k, double& s0)
{
    if (VAR) {
        double S[2][2][2];
        gather_sigma_g<true>(g, i, j, k, S);
        return stencil27(g, S, P, s0);
    }
    return stencil27_c(g, g.csig, P, s0);
}

// one Gauss-Seidel sweep A -> B (A: previous sweep)
template <bool VAR>
__device__ void tail_sweep(const Lev& g, int tz, const double* __restrict__ rhs, const double* __restrict__ A, double* __restrict__ B)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const int nodes = g.nn[0] * g.nn[1] * g.nn[2];
    for (int t = tid; t < nodes; t += nt) B[t] = A[t];
    __syncthreads();
    const int nch = (g.nn[2] + tz - 1) / tz;
    for (int p = 0; p < tz; ++p)
        for (int col = 0; col < 4; ++col) {
            const int ci = col & 1, cj = col >> 1;
            const int ni = (g.nn[0] - ci + 1) / 2, nj = (g.nn[1] - cj + 1) / 2;
            const int total = ni * nj * nch;
            for (int t = tid; t < total; t += nt) {
                const int ii = t % ni, r = t / ni, jj = r % nj, c = r / nj;
                const int i = 2 * ii + ci, j = 2 * jj + cj, k = c * tz + p;
                if (k >= g.nn[2]) continue;
                const int id = (k * g.nn[1] + j) * g.nn[0] + i;
                if (node_masked(g, i, j, k)) { B[id] = 0.0; continue; }
                double P[3][3][3], s0;
                tail_gather(g, tz, B, A, i, j, k, P);
                const double Ax = tail_Lphi<VAR>(g, P, i, j, k, s0);
                B[id] = P[1][1][1] + (rhs[id] - Ax) / s0;
            }
            __syncthreads();
        }
}

// out = rhs - L x
template <bool VAR>
__device__ void tail_residual(const Lev& g, const double* __restrict__ x, const double* __restrict__ rhs, double* __restrict__ out)
{
    const int nxy = g.nn[0] * g.nn[1], nodes = nxy * g.nn[2];
    for (int t = threadIdx.x; t < nodes; t += blockDim.x) {
        const int k = t / nxy, r = t - k * nxy, j = r / g.nn[0], i = r - j * g.nn[0];
        double v = 0.0;
        if (!node_masked(g, i, j, k)) {
            double P[3][3][3], s0;
            tail_gather(g, 1 << 20, x, x, i, j, k, P);
            v = rhs[t] - tail_Lphi<VAR>(g, P, i, j, k, s0);
        }
        out[t] = v;
    }
    __syncthreads();
}

// crse = full weighting of fine (A.5)
__device__ void tail_restrict(const Lev& F, const Lev& C, const double* __restrict__ fine, double* __restrict__ crse)
{
    const int nxy = C.nn[0] * C.nn[1], nodes = nxy * C.nn[2];
    for (int t = threadIdx.x; t < nodes; t += blockDim.x) {
        const int k = t / nxy, r = t - k * nxy, j = r / C.nn[0], i = r - j * C.nn[0];
        double s = 0.0;
        if (!node_masked(C, i, j, k)) {
#pragma unroll
            for (int c = -1; c <= 1; ++c)
#pragma unroll
                for (int b = -1; b <= 1; ++b)
#pragma unroll
                    for (int a = -1; a <= 1; ++a) {
                        const double w = (a ? 1.0 : 2.0) * (b ? 1.0 : 2.0) * (c ? 1.0 : 2.0);
                        s += w * fine[(nmap(2 * k + c, F.n[2], F.per[2]) * F.nn[1] + nmap(2 * j + b, F.n[1], F.per[1])) * F.nn[0] +
                                      nmap(2 * i + a, F.n[0], F.per[0])];
                    }
            s *= 1.0 / 64.0;
        }
        crse[t] = s;
    }
    __syncthreads();
}

// fine += P crse (A.6)
template <bool VAR>
__device__ void tail_interp_add(const Lev& F, const Lev& C, double* __restrict__ fine, const double* __restrict__ crse)
{
    const int nxy = F.nn[0] * F.nn[1], nodes = nxy * F.nn[2];
    Interp<VAR> x{F, C, crse, 0};
    for (int t = threadIdx.x; t < nodes; t += blockDim.x) {
        const int k = t / nxy, r = t - k * nxy, j = r / F.nn[0], i = r - j * F.nn[0];
        if (node_masked(F, i, j, k)) continue;
        fine[t] += x.value(i, j, k, k);
    }
    __syncthreads();
}

template <bool VAR>
__global__ void __launch_bounds__(TAIL_THREADS, 1) k_coarse_tail(const TailPlan P)
{
    extern __shared__ __align__(16) double tail_sm[];
    __shared__ Lev LV[TAIL_MAX_LEV];
    __shared__ double sh[34];
    const int tid = threadIdx.x, nt = blockDim.x;
    pdl_trigger();
    if (tid < P.nlev) {
        Lev g = P.lv[tid].g;
        g.sigma = VAR ? tail_sm + P.lv[tid].s : nullptr;
        LV[tid] = g;
    }
    // sigma of every tail level (constant during a solve; a few KB)
    if (VAR)
        for (int l = 0; l < P.nlev; ++l) {
            const TailLev& T = P.lv[l];
            const int n0 = T.g.n[0], n01 = n0 * T.g.n[1], cells = n01 * T.g.n[2];
            for (int t = tid; t < cells; t += nt) {
                const int k = t / n01, r = t - k * n01, j = r / n0, i = r - j * n0;
                tail_sm[T.s + t] = T.sigma_g[k * T.cps_g + (long long)j * T.cpx_g + i];
            }
        }
    pdl_wait();   // res of the first tail level comes from the restriction kernel before us
    {
        const TailLev& T = P.lv[0];
        const int n0 = T.g.nn[0], n01 = n0 * T.g.nn[1], nodes = n01 * T.g.nn[2];
        for (int t = tid; t < nodes; t += nt) {
            const int k = t / n01, r = t - k * n01, j = r / n0, i = r - j * n0;
            tail_sm[T.r + t] = P.res_in[k * P.ps_io + (long long)j * P.px_io + i];
        }
    }
    __syncthreads();
    // ---- down-leg ----
    for (int l = 0; l + 1 < P.nlev; ++l) {
        const TailLev& T = P.lv[l];
        const Lev& g = LV[l];
        const int nodes = g.nn[0] * g.nn[1] * g.nn[2];
        double *x = tail_sm + T.a, *y = tail_sm + T.b;
        const double* rhs = tail_sm + T.r;
        for (int t = tid; t < nodes; t += nt) x[t] = 0.0;
        __syncthreads();
        for (int s = 0; s < P.nu1 * P.nsw; ++s) { tail_sweep<VAR>(g, T.tz, rhs, x, y); double* q = x; x = y; y = q; }
        tail_residual<VAR>(g, x, rhs, y);                                 // rescor in the free ping-pong array
        tail_restrict(g, LV[l + 1], y, tail_sm + P.lv[l + 1].r);
        if (x != tail_sm + T.a) {   // odd sweep count: keep the correction in `a`
            for (int t = tid; t < nodes; t += nt) tail_sm[T.a + t] = x[t];
            __syncthreads();
        }
    }
    // ---- bottom ----
    {
        const TailLev& T = P.lv[P.nlev - 1];
        bottom_bicgstab_body<VAR>(LV[P.nlev - 1], tail_sm + T.a, tail_sm + T.r, tail_sm + P.work, P.maxiter, P.rtol, P.atol, P.singular,
                                  P.nsw, P.bottom_solver, P.info, sh);
        __syncthreads();
    }
    // ---- up-leg ----
    for (int l = P.nlev - 2; l >= 0; --l) {
        const TailLev& T = P.lv[l];
        const Lev& g = LV[l];
        const int nodes = g.nn[0] * g.nn[1] * g.nn[2];
        double *x = tail_sm + T.a, *y = tail_sm + T.b;
        tail_interp_add<VAR>(g, LV[l + 1], x, tail_sm + P.lv[l + 1].a);
        for (int s = 0; s < P.nu2 * P.nsw; ++s) { tail_sweep<VAR>(g, T.tz, tail_sm + T.r, x, y); double* q = x; x = y; y = q; }
        if (x != tail_sm + T.a) {
            for (int t = tid; t < nodes; t += nt) tail_sm[T.a + t] = x[t];
            __syncthreads();
        }
    }
    {
        const TailLev& T = P.lv[0];
        const int n0 = T.g.nn[0], n01 = n0 * T.g.nn[1], nodes = n01 * T.g.nn[2];
        for (int t = tid; t < nodes; t += nt) {
            const int k = t / n01, r = t - k * n01, j = r / n0, i = r - j * n0;
            P.cor_out[k * P.ps_io + (long long)j * P.px_io + i] = tail_sm[T.a + t];
        }
    }
}

}  // namespace b200np_dev
