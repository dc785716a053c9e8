// b200eb.cu -- the EB (embedded boundary, cut cell) nodal projection behind the C ABI b200eb_* (include/b200np.h):
// Hydro::NodalProjector over amrex::MLMG / MLNodeLaplacian built with an EBFArrayBoxFactory, as incflo drives it under
// AMREX_USE_EB (src/projection/incflo_apply_nodal_projection.cpp:130-136 set_eb_*, :181-201 projector + setEBInflowVelocity,
// :215-266 project / copy-out), BASELINE configs[4] (test_3d/benchmark.channel_cylinder-x).
//
// Discretisation (restated in oracle/eb_oracle.py, which tests/ compare this file with; derivation in its header):
//   L(a, b) = - sum_cells sigma_c sum_d dxinv_d^2 int_{fluid part of c} d_d N_a d_d N_b     (Q1 stiffness over the fluid only;
//             the integrals are combinations of the volume fraction and the 18 monomial integrals MLNodeLaplacian::buildIntegral
//             stores -- mlndlap_set_connection / mlndlap_set_stencil_eb)
//   rhs(a)  = - sum_cells u_c . int_fluid grad N_a  (+ sum_cells (u_eb . n)_c int_{EB face} N_a dA)            (mlndlap_divu_eb)
//   u_c    -= sigma_c (1/V_c) int_fluid grad phi;  grad phi = that average; covered cells: 0                   (mlndlap_mknewu_eb)
// Multigrid = AMReX's "RAP" strategy (what MLNodeLaplacian switches to with EB): every level is a symmetric 27-point stencil,
// coarse levels are Galerkin products (1/8) P^T A P with trilinear P; smoother = Gauss-Seidel in 8 colours
// (mlndlap_gscolor_sten), nodes with a zero diagonal (covered, Dirichlet) hold 0; MLMG V-cycle, BiCGStab bottom solve in one CTA.
//
// Layout: every nodal array of a level is stored COLOUR-MAJOR -- the 8 colours c = (i&1) + 2(j&1) + 4(k&1) one after the other, each a
// dense (k/2, j/2, i/2) block -- so that a colour sweep reads its own 27 coefficients and writes its nodes with unit stride, and
// the neighbours it reads (always other colours) are unit-stride runs of other blocks.  A row is stored completely (27 arrays per
// level, t = (di+1) + 3(dj+1) + 9(dk+1)): 216 B of coefficients per node and sweep is the traffic that bounds the smoother.
// One AMR level, one box, one GPU: the first correct path of this operator.
#include "../../include/b200np.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

#define ECK(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            fprintf(stderr, "b200eb: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            throw int(B200NP_ERR_CUDA);                                                             \
        }                                                                                           \
    } while (0)

struct EbLev {
    int n[3], nn[3], per[3];
    int dirlo[3], dirhi[3];   // Dirichlet faces: their nodes are masked
    int H[3];                 // extent of a colour block per direction: (nn + 1) / 2 (the odd-parity block of an odd extent has one
    int CS;                   // unused slot per row: a hole; holes hold zeros everywhere and are never written); CS = H[0] H[1] H[2]
    long long nnode;          // allocated length of a nodal array = 8 CS (< 2^31)
    double* st;               // 27 coefficient arrays, st + t * nnode; t = 13: diagonal
    const double* sigma;         // level 0 with variable sigma: the cell array (natural order) the flag-2 rows are computed from; else nullptr
    double hinv2_12;             // 1 / (12 h^2) of the level (isotropic cells)
    const unsigned char* flag;   // 2: uncut neighbourhood with VARIABLE sigma (level 0 only): the row follows from the 8 sigmas around the node
                                 // (corners sigma_c / 12h^2, edges (sigma_a + sigma_b) / 12h^2, faces 0, diagonal -sum / 3h^2), not read from st;
                                 // 1: the row of this node is the canonical row of an uncut neighbourhood with constant sigma (faces 0, edges
    const double* canon;         // canon[0], corners canon[1], diagonal canon[2]): the kernels do not read its 27 coefficients.  All 0 with
};                               // variable sigma.  canon lives in device memory: a captured V-cycle graph must see the current sigma.

// Programmatic dependent launch: pdl_wait() blocks until the predecessor kernel has completed and its writes are visible (must precede
// the first access to data it produced); pdl_trigger() lets the successor become resident early (it waits in its own pdl_wait()).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// caller array with its own box, component stride
struct EFab {
    double* p;
    int lo[3];
    int nx, ny;
    long long cs;
    __host__ __device__ __forceinline__ long long idx(int i, int j, int k) const
    {
        return (i - lo[0]) + (long long)nx * ((j - lo[1]) + (long long)ny * (k - lo[2]));
    }
};

// position of node (i, j, k): colour block ((k&1) 4 + (j&1) 2 + (i&1)), then (k/2, j/2, i/2) -- a sum of one term per direction
__device__ __forceinline__ int nidx(const EbLev& L, int i, int j, int k)
{
    return ((k & 1) * 4 + (j & 1) * 2 + (i & 1)) * L.CS + ((k >> 1) * L.H[1] + (j >> 1)) * L.H[0] + (i >> 1);
}
// position t -> node (i, j, k); false: t is a hole
__device__ __forceinline__ bool ndecode(const EbLev& L, long long tt, int& i, int& j, int& k)
{
    const int t = (int)tt;
    const int c = t / L.CS, l = t - c * L.CS;
    const int r = l / L.H[0];
    i = 2 * (l - r * L.H[0]) + (c & 1);
    const int k2 = r / L.H[1];
    j = 2 * (r - k2 * L.H[1]) + ((c >> 1) & 1);
    k = 2 * k2 + (c >> 2);
    return i < L.nn[0] && j < L.nn[1] && k < L.nn[2];
}
// the same for the 26 neighbours: index = X[di] + Y[dj] + Z[dk] (periodic wrap; clamped where the neighbour does not exist -- its
// coefficient is exactly zero)
struct NbIdx {
    int X[3], Y[3], Z[3];
};
__device__ __forceinline__ void nb_index(const EbLev& L, int i, int j, int k, NbIdx& q)
{
    const int p[3] = {i, j, k};
    int c[3][3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        int lo = p[d] - 1, hi = p[d] + 1;
        if (lo < 0) lo = L.per[d] ? L.nn[d] - 1 : p[d];
        if (hi >= L.nn[d]) hi = L.per[d] ? 0 : p[d];
        c[d][0] = lo; c[d][1] = p[d]; c[d][2] = hi;
    }
    const int sy = L.H[0], sz = L.H[0] * L.H[1];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        q.X[t] = (c[0][t] & 1) * L.CS + (c[0][t] >> 1);
        q.Y[t] = (c[1][t] & 1) * 2 * L.CS + (c[1][t] >> 1) * sy;
        q.Z[t] = (c[2][t] & 1) * 4 * L.CS + (c[2][t] >> 1) * sz;
    }
}
// coordinates of q - 1, q, q + 1 per direction: periodic wrap; outside a non-periodic domain the coordinate is clamped (the
// coefficient towards it is exactly zero by construction) and flagged
struct Nb {
    int c[3][3];
    bool ok[3][3];
};
__device__ __forceinline__ void nb_coords(const EbLev& L, int i, int j, int k, Nb& q)
{
    const int p[3] = {i, j, k};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        q.c[d][1] = p[d]; q.ok[d][1] = true;
        int lo = p[d] - 1, hi = p[d] + 1;
        bool oklo = true, okhi = true;
        if (lo < 0) { if (L.per[d]) lo = L.nn[d] - 1; else { lo = p[d]; oklo = false; } }
        if (hi >= L.nn[d]) { if (L.per[d]) hi = 0; else { hi = p[d]; okhi = false; } }
        q.c[d][0] = lo; q.ok[d][0] = oklo;
        q.c[d][2] = hi; q.ok[d][2] = okhi;
    }
}
__device__ __forceinline__ bool node_dirichlet(const EbLev& L, int i, int j, int k)
{
    const int p[3] = {i, j, k};
    bool m = false;
#pragma unroll
    for (int d = 0; d < 3; ++d)
        if (!L.per[d]) m = m || (L.dirlo[d] && p[d] == 0) || (L.dirhi[d] && p[d] == L.nn[d] - 1);
    return m;
}

__device__ __forceinline__ double canonical_entry(const EbLev& L, int tt)
{
    const int nz = (tt % 3 != 1) + ((tt / 3) % 3 != 1) + (tt / 9 != 1);
    return nz == 0 ? L.canon[2] : nz == 1 ? 0.0 : nz == 2 ? L.canon[0] : L.canon[1];
}

template <bool MAXR>
__device__ __forceinline__ double eb_block_reduce(double v, double* sh)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double w = __shfl_xor_sync(0xffffffffu, v, o);
        v = MAXR ? fmax(v, w) : v + w;
    }
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5, nw = (blockDim.x * blockDim.y + 31) >> 5;
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    if (wid == 0) {
        v = lane < nw ? sh[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double w = __shfl_xor_sync(0xffffffffu, v, o);
            v = MAXR ? fmax(v, w) : v + w;
        }
        if (lane == 0) sh[32] = v;
    }
    __syncthreads();
    return sh[32];
}

// ---- cut-cell integrals --------------------------------------------------------------------------------------------------
// geo: 19 cell arrays (natural cell order): 0 = volume fraction, 1 + m = the m-th monomial integral (amrex i_S_* order:
// x y z x2 y2 z2 xy xz yz x2y x2z xy2 y2z xz2 yz2 x2y2 x2z2 y2z2)
__device__ __forceinline__ int mom_index(int px, int py, int pz)
{
    // -1: the volume itself
    constexpr int T[3][3][3] = {   // [px][py][pz]
        {{-1, 2, 5}, {1, 8, 14}, {4, 12, 17}},
        {{0, 7, 13}, {6, -2, -2}, {11, -2, -2}},
        {{3, 10, 16}, {9, -2, -2}, {15, -2, -2}}};
    return T[px][py][pz];
}
// M[d][p][q] = int x_e^p x_f^q over the fluid part, (e, f) = the two directions other than d
__device__ __forceinline__ void load_moments(const double* __restrict__ geo, long long ncell, long long c, double M[3][3][3])
{
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int px = d == 0 ? 0 : p, py = d == 0 ? p : (d == 1 ? 0 : q), pz = d == 2 ? 0 : q;
                const int m = mom_index(px, py, pz);
                M[d][p][q] = geo[(long long)(m + 1) * ncell + c];
            }
}
// int_F d_d N_a,  N_a = prod (1/2 + s x)
__device__ __forceinline__ double grad_integral(const double M[3][3][3], int d, int a)
{
    const int e = d == 0 ? 1 : 0, f = d == 2 ? 1 : 2;
    const double sd = ((a >> d) & 1) ? 1.0 : -1.0, se = ((a >> e) & 1) ? 1.0 : -1.0, sf = ((a >> f) & 1) ? 1.0 : -1.0;
    return sd * (0.25 * M[d][0][0] + 0.5 * se * M[d][1][0] + 0.5 * sf * M[d][0][1] + se * sf * M[d][1][1]);
}
// - sum_d dxinv_d^2 int_F d_d N_a d_d N_b
__device__ __forceinline__ double stiff_entry(const double M[3][3][3], const double dh[3], int a, int b)
{
    double tot = 0.0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int e = d == 0 ? 1 : 0, f = d == 2 ? 1 : 2;
        const double sad = ((a >> d) & 1) ? 1.0 : -1.0, sbd = ((b >> d) & 1) ? 1.0 : -1.0;
        const double sae = ((a >> e) & 1) ? 1.0 : -1.0, sbe = ((b >> e) & 1) ? 1.0 : -1.0;
        const double saf = ((a >> f) & 1) ? 1.0 : -1.0, sbf = ((b >> f) & 1) ? 1.0 : -1.0;
        const double ce[3] = {0.25, 0.5 * (sae + sbe), sae * sbe};
        const double cf[3] = {0.25, 0.5 * (saf + sbf), saf * sbf};
        double acc = 0.0;
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int q = 0; q < 3; ++q) acc += ce[p] * cf[q] * M[d][p][q];
        tot += dh[d] * sad * sbd * acc;
    }
    return -tot;
}

// ---- set-up kernels ---------------------------------------------------------------------------------------------------------
// caller's vfrac / intg arrays (own boxes) -> geo
__global__ void __launch_bounds__(256) k_eb_copy_geo(int nx, int ny, int nz, EFab vf, EFab ig, double* __restrict__ geo)
{
    const long long N = (long long)nx * ny * nz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < N; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % nx), j = (int)((t / nx) % ny), k = (int)(t / ((long long)nx * ny));
        geo[t] = vf.p[vf.idx(i, j, k)];
        const long long q = ig.idx(i, j, k);
#pragma unroll
        for (int m = 0; m < 18; ++m) geo[(long long)(m + 1) * N + t] = ig.p[m * ig.cs + q];
    }
}
// optional EB-inflow data: vn = u_eb . n and the 8 surface integrals (B_1 B_x B_y B_z B_xy B_xz B_yz B_xyz) -> ebf (9 cell arrays)
__global__ void __launch_bounds__(256) k_eb_copy_flow(int nx, int ny, int nz, EFab ev, EFab bn, EFab bi, double* __restrict__ ebf)
{
    const long long N = (long long)nx * ny * nz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < N; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % nx), j = (int)((t / nx) % ny), k = (int)(t / ((long long)nx * ny));
        const long long qe = ev.idx(i, j, k), qn = bn.idx(i, j, k), qb = bi.idx(i, j, k);
        ebf[t] = ev.p[qe] * bn.p[qn] + ev.p[ev.cs + qe] * bn.p[bn.cs + qn] + ev.p[2 * ev.cs + qe] * bn.p[2 * bn.cs + qn];
#pragma unroll
        for (int m = 0; m < 8; ++m) ebf[(long long)(m + 1) * N + t] = bi.p[m * bi.cs + qb];
    }
}
// sigma: caller array or constant -> dense cell array
__global__ void __launch_bounds__(256) k_eb_copy_sigma(int nx, int ny, int nz, EFab sg, double cs, double* __restrict__ out)
{
    const long long N = (long long)nx * ny * nz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < N; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % nx), j = (int)((t / nx) % ny), k = (int)(t / ((long long)nx * ny));
        out[t] = sg.p ? sg.p[sg.idx(i, j, k)] : cs;
    }
}

// level-0 stencil: the complete row of every node from sigma and the cut-cell integrals of its (up to) 8 cells
__global__ void __launch_bounds__(128) k_eb_stencil0(const EbLev L, const double* __restrict__ geo, const double* __restrict__ sigma,
                                                     double dhx, double dhy, double dhz)
{
    const long long ncell = (long long)L.n[0] * L.n[1] * L.n[2];
    const double dh[3] = {dhx, dhy, dhz};
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < L.nnode; t += (long long)gridDim.x * blockDim.x) {
        int i, j, k;
        if (!ndecode(L, t, i, j, k)) continue;
        if (L.flag[t] == 1) {
#pragma unroll
            for (int tt = 0; tt < 27; ++tt) L.st[(long long)tt * L.nnode + t] = canonical_entry(L, tt);
            continue;
        }
        double row[27];
#pragma unroll
        for (int q = 0; q < 27; ++q) row[q] = 0.0;
        const int p[3] = {i, j, k};
        if (!node_dirichlet(L, i, j, k)) {
#pragma unroll
            for (int a = 0; a < 8; ++a) {   // the node is corner a of cell c = node - a
                int c[3];
                bool ok = true;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    c[d] = p[d] - ((a >> d) & 1);
                    if (c[d] < 0) { if (L.per[d]) c[d] = L.n[d] - 1; else ok = false; }
                    if (c[d] >= L.n[d]) ok = false;   // only a non-periodic top node (periodic nodes end at n - 1)
                }
                if (!ok) continue;
                const long long cc = ((long long)c[2] * L.n[1] + c[1]) * L.n[0] + c[0];
                if (geo[cc] == 0.0) continue;   // covered cell
                double M[3][3][3];
                load_moments(geo, ncell, cc, M);
                const double sg = sigma[cc];
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const int tt = (((b & 1) - (a & 1)) + 1) + 3 * ((((b >> 1) & 1) - ((a >> 1) & 1)) + 1) + 9 * (((b >> 2) - (a >> 2)) + 1);
                    row[tt] += sg * stiff_entry(M, dh, a, b);
                }
            }
            Nb q;
            nb_coords(L, i, j, k, q);
#pragma unroll
            for (int tt = 0; tt < 27; ++tt) {
                const int di = tt % 3, dj = (tt / 3) % 3, dk = tt / 9;
                if (tt == 13) continue;
                if (!(q.ok[0][di] && q.ok[1][dj] && q.ok[2][dk]) || node_dirichlet(L, q.c[0][di], q.c[1][dj], q.c[2][dk])) row[tt] = 0.0;
            }
        }
#pragma unroll
        for (int tt = 0; tt < 27; ++tt) L.st[(long long)tt * L.nnode + t] = row[tt];
    }
}

// level 0: a node has the canonical row iff its 8 cells exist and are uncut and neither it nor a neighbour is a Dirichlet node
__global__ void __launch_bounds__(256) k_eb_flag0(const EbLev L, const double* __restrict__ geo, unsigned char* __restrict__ flag, int kind)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < L.nnode; t += (long long)gridDim.x * blockDim.x) {
        int i, j, k;
        if (!ndecode(L, t, i, j, k)) continue;
        const int p[3] = {i, j, k};
        bool reg = true;
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            int c[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                c[d] = p[d] - ((a >> d) & 1);
                if (c[d] < 0) { if (L.per[d]) c[d] = L.n[d] - 1; else reg = false; }
                if (c[d] >= L.n[d]) reg = false;
            }
            if (reg) reg = geo[((long long)c[2] * L.n[1] + c[1]) * L.n[0] + c[0]] == 1.0;
        }
        if (reg) {
            Nb q;
            nb_coords(L, i, j, k, q);
#pragma unroll
            for (int tt = 0; tt < 27; ++tt) reg = reg && !node_dirichlet(L, q.c[0][tt % 3], q.c[1][(tt / 3) % 3], q.c[2][tt / 9]);
        }
        flag[t] = reg ? kind : 0;
    }
}
// coarser level: canonical iff the 27 fine nodes under it are, and no neighbour is a Dirichlet node
__global__ void __launch_bounds__(256) k_eb_flag_coarse(const EbLev C, const EbLev F, unsigned char* __restrict__ flag)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < C.nnode; t += (long long)gridDim.x * blockDim.x) {
        int I, J, K;
        if (!ndecode(C, t, I, J, K)) continue;
        Nb q;
        nb_coords(F, 2 * I, 2 * J, 2 * K, q);
        bool reg = true;
#pragma unroll
        for (int tt = 0; tt < 27; ++tt) {
            const int di = tt % 3, dj = (tt / 3) % 3, dk = tt / 9;
            reg = reg && q.ok[0][di] && q.ok[1][dj] && q.ok[2][dk] && F.flag[nidx(F, q.c[0][di], q.c[1][dj], q.c[2][dk])] == 1;
        }
        if (reg) {
            Nb qc;
            nb_coords(C, I, J, K, qc);
#pragma unroll
            for (int tt = 0; tt < 27; ++tt) {
                const int di = tt % 3, dj = (tt / 3) % 3, dk = tt / 9;
                reg = reg && qc.ok[0][di] && qc.ok[1][dj] && qc.ok[2][dk] && !node_dirichlet(C, qc.c[0][di], qc.c[1][dj], qc.c[2][dk]);
            }
        }
        flag[t] = reg ? 1 : 0;
    }
}
// canonical row of every level for a constant sigma: level l has cell size h 2^l
__global__ void k_eb_set_canon(double* canon, int nlev, double sigma, double hinv2)
{
    const int l = threadIdx.x;
    if (l >= nlev) return;
    double s = sigma * hinv2;
    for (int q = 0; q < l; ++q) s *= 0.25;
    canon[3 * l + 0] = s / 6.0;
    canon[3 * l + 1] = s / 12.0;
    canon[3 * l + 2] = -8.0 * s / 3.0;
}

__device__ __forceinline__ double w1(int t) { return t == 0 ? 1.0 : 0.5; }

// Galerkin coarse operator: A_c(I, I + D) = (1/8) sum_{a, o} w(a) A_f(2I + a, 2I + a + o) w(a + o - 2D), complete rows
__global__ void __launch_bounds__(128) k_eb_rap(const EbLev C, const EbLev F)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < C.nnode; t += (long long)gridDim.x * blockDim.x) {
        if (C.flag[t] == 1) {
            for (int tt = 0; tt < 27; ++tt) C.st[(long long)tt * C.nnode + t] = canonical_entry(C, tt);
            continue;
        }
        int I, J, K;
        if (!ndecode(C, t, I, J, K)) continue;
        double acc[27];
#pragma unroll
        for (int q = 0; q < 27; ++q) acc[q] = 0.0;
        const bool masked = node_dirichlet(C, I, J, K);
        if (!masked) {
            for (int az = -1; az <= 1; ++az)
                for (int ay = -1; ay <= 1; ++ay)
                    for (int ax = -1; ax <= 1; ++ax) {
                        int f[3] = {2 * I + ax, 2 * J + ay, 2 * K + az};
                        bool ok = true;
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            if (f[d] < 0) { if (F.per[d]) f[d] += F.nn[d]; else ok = false; }
                            if (f[d] >= F.nn[d]) { if (F.per[d]) f[d] -= F.nn[d]; else ok = false; }
                        }
                        if (!ok) continue;
                        const long long pf = nidx(F, f[0], f[1], f[2]);
                        const double wa = 0.125 * w1(ax) * w1(ay) * w1(az);
                        for (int o = 0; o < 27; ++o) {
                            const double A = F.st[(long long)o * F.nnode + pf];
                            if (A == 0.0) continue;
                            const int gx = ax + o % 3 - 1, gy = ay + (o / 3) % 3 - 1, gz = az + o / 9 - 1;   // g - 2I
                            // coarse targets D with |g - 2D| <= 1: g even -> D = g / 2 (weight 1), g odd -> (g - 1) / 2 and (g + 1) / 2 (weight 1/2)
                            const int x0 = (gx - 1 + 4) / 2 - 2 + ((gx & 1) ? 0 : 1), nxo = (gx & 1) ? 2 : 1;
                            const int y0 = (gy - 1 + 4) / 2 - 2 + ((gy & 1) ? 0 : 1), nyo = (gy & 1) ? 2 : 1;
                            const int z0 = (gz - 1 + 4) / 2 - 2 + ((gz & 1) ? 0 : 1), nzo = (gz & 1) ? 2 : 1;
                            const double wo = wa * A * ((gx & 1) ? 0.5 : 1.0) * ((gy & 1) ? 0.5 : 1.0) * ((gz & 1) ? 0.5 : 1.0);
                            for (int dz = 0; dz < nzo; ++dz)
                                for (int dy = 0; dy < nyo; ++dy)
                                    for (int dx = 0; dx < nxo; ++dx) acc[(x0 + dx + 1) + 3 * (y0 + dy + 1) + 9 * (z0 + dz + 1)] += wo;
                        }
                    }
            Nb q;
            nb_coords(C, I, J, K, q);
            for (int tt = 0; tt < 27; ++tt) {
                const int di = tt % 3, dj = (tt / 3) % 3, dk = tt / 9;
                if (tt == 13) continue;
                if (!(q.ok[0][di] && q.ok[1][dj] && q.ok[2][dk]) || node_dirichlet(C, q.c[0][di], q.c[1][dj], q.c[2][dk])) acc[tt] = 0.0;
            }
        }
        for (int tt = 0; tt < 27; ++tt) C.st[(long long)tt * C.nnode + t] = acc[tt];
    }
}

// ---- multigrid kernels ------------------------------------------------------------------------------------------------------
// loads the compiler may not sink towards their use: with one node per thread the kernels below are pure latency unless every load
// of a node is in flight before the first multiply (nvcc otherwise pairs each load with its FMA: 26 dependent round trips per node)
__device__ __forceinline__ double ld_early(const double* p)
{
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_early_nc(const double* p)
{
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
// BATCH = true: every load of a node is issued before the first multiply (~120 registers, few resident warps): right for the
// coarser levels, whose launches are a handful of waves and pure latency.  BATCH = false: plain loads the compiler pairs with
// their FMAs (~40 registers, full occupancy): measured faster on a level of millions of nodes, where many resident warps hide the latency.
// sum over the 26 neighbours of A(p, q) x(q)
template <bool BATCH>
__device__ __forceinline__ double offdiag_sum(const EbLev& L, long long p, int i, int j, int k, const double* x)
{
    NbIdx q;
    nb_index(L, i, j, k, q);
    double ax = 0.0;
    if constexpr (!BATCH) {
#pragma unroll
        for (int tt = 0; tt < 27; ++tt) {
            if (tt == 13) continue;
            ax += __ldg(L.st + (long long)tt * L.nnode + p) * x[q.X[tt % 3] + q.Y[(tt / 3) % 3] + q.Z[tt / 9]];
        }
        return ax;
    } else {
        double c[27], v[27];
#pragma unroll
        for (int tt = 0; tt < 27; ++tt) {
            if (tt == 13) continue;
            c[tt] = ld_early_nc(L.st + (long long)tt * L.nnode + p);
            v[tt] = ld_early(x + (q.X[tt % 3] + q.Y[(tt / 3) % 3] + q.Z[tt / 9]));
        }
#pragma unroll
        for (int tt = 26; tt >= 0; --tt) {   // the first product needs the LAST loads: nothing can stall before all of them are issued
            if (tt == 13) continue;
            ax += c[tt] * v[tt];
        }
        return ax;
    }
}
// the same sum for a node with the canonical row: 12 edge and 8 corner neighbours, no coefficient loads
template <bool BATCH>
__device__ __forceinline__ double offdiag_sum_regular(const EbLev& L, int i, int j, int k, const double* x)
{
    NbIdx q;
    nb_index(L, i, j, k, q);
    double se = 0.0, sc = 0.0;
    if constexpr (!BATCH) {
#pragma unroll
        for (int tt = 0; tt < 27; ++tt) {
            const int nz = (tt % 3 != 1) + ((tt / 3) % 3 != 1) + (tt / 9 != 1);
            if (nz == 2) se += x[q.X[tt % 3] + q.Y[(tt / 3) % 3] + q.Z[tt / 9]];
            if (nz == 3) sc += x[q.X[tt % 3] + q.Y[(tt / 3) % 3] + q.Z[tt / 9]];
        }
        return L.canon[0] * se + L.canon[1] * sc;
    } else {
        double v[27];
#pragma unroll
        for (int tt = 0; tt < 27; ++tt) {
            const int nz = (tt % 3 != 1) + ((tt / 3) % 3 != 1) + (tt / 9 != 1);
            if (nz >= 2) v[tt] = ld_early(x + (q.X[tt % 3] + q.Y[(tt / 3) % 3] + q.Z[tt / 9]));
        }
#pragma unroll
        for (int tt = 26; tt >= 0; --tt) {
            const int nz = (tt % 3 != 1) + ((tt / 3) % 3 != 1) + (tt / 9 != 1);
            if (nz == 2) se += v[tt];
            if (nz == 3) sc += v[tt];
        }
        return L.canon[0] * se + L.canon[1] * sc;
    }
}

// Interior nodes (no wrap, no clamp: 0 < i < nn - 1 in every direction): the 26 neighbours sit at p + a constant that depends on the
// colour only -- stepping down from an even coordinate goes to the odd block one slot earlier, from an odd one to the even block at the
// same slot, and so on.  D[d][0 / 2] = offset of the lower / upper neighbour in direction d; D[d][1] = 0.
struct NbOff {
    int D[3][3];
};
__device__ __forceinline__ NbOff nb_offsets(const EbLev& L, int color)
{
    NbOff o;
    const int cs[3] = {L.CS, 2 * L.CS, 4 * L.CS}, st[3] = {1, L.H[0], L.H[0] * L.H[1]};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const bool odd = (color >> d) & 1;
        o.D[d][0] = odd ? -cs[d] : cs[d] - st[d];
        o.D[d][1] = 0;
        o.D[d][2] = odd ? -cs[d] + st[d] : cs[d];
    }
    return o;
}
template <bool BATCH>
__device__ __forceinline__ double offdiag_sum_inner(const EbLev& L, int p, const NbOff& o, const double* x)
{
    const double* xp = x + p;
    const double* cp = L.st + p;
    double ax = 0.0;
    if constexpr (!BATCH) {
#pragma unroll
        for (int tt = 0; tt < 27; ++tt) {
            if (tt == 13) continue;
            ax += __ldg(cp + (long long)tt * L.nnode) * xp[o.D[0][tt % 3] + o.D[1][(tt / 3) % 3] + o.D[2][tt / 9]];
        }
        return ax;
    } else {
        double c[27], v[27];
#pragma unroll
        for (int tt = 0; tt < 27; ++tt) {
            if (tt == 13) continue;
            c[tt] = ld_early_nc(cp + (long long)tt * L.nnode);
            v[tt] = ld_early(xp + (o.D[0][tt % 3] + o.D[1][(tt / 3) % 3] + o.D[2][tt / 9]));
        }
#pragma unroll
        for (int tt = 26; tt >= 0; --tt) {
            if (tt == 13) continue;
            ax += c[tt] * v[tt];
        }
        return ax;
    }
}
template <bool BATCH>
__device__ __forceinline__ double offdiag_sum_regular_inner(const EbLev& L, int p, const NbOff& o, const double* x)
{
    const double* xp = x + p;
    double se = 0.0, sc = 0.0;
    if constexpr (!BATCH) {
#pragma unroll
        for (int tt = 0; tt < 27; ++tt) {
            const int nz = (tt % 3 != 1) + ((tt / 3) % 3 != 1) + (tt / 9 != 1);
            if (nz == 2) se += xp[o.D[0][tt % 3] + o.D[1][(tt / 3) % 3] + o.D[2][tt / 9]];
            if (nz == 3) sc += xp[o.D[0][tt % 3] + o.D[1][(tt / 3) % 3] + o.D[2][tt / 9]];
        }
        return L.canon[0] * se + L.canon[1] * sc;
    } else {
        double v[27];
#pragma unroll
        for (int tt = 0; tt < 27; ++tt) {
            const int nz = (tt % 3 != 1) + ((tt / 3) % 3 != 1) + (tt / 9 != 1);
            if (nz >= 2) v[tt] = ld_early(xp + (o.D[0][tt % 3] + o.D[1][(tt / 3) % 3] + o.D[2][tt / 9]));
        }
#pragma unroll
        for (int tt = 26; tt >= 0; --tt) {
            const int nz = (tt % 3 != 1) + ((tt / 3) % 3 != 1) + (tt / 9 != 1);
            if (nz == 2) se += v[tt];
            if (nz == 3) sc += v[tt];
        }
        return L.canon[0] * se + L.canon[1] * sc;
    }
}
// sum_q!=p A(p, q) x(q) and the diagonal of node p = (i, j, k) of colour `color`, by the cheapest applicable path; false: inactive node
// row of an interior node with 8 uncut cells from their sigmas (SURVEY A.3 with isotropic cells); all loads before the first multiply
__device__ __forceinline__ void sigma_row(const EbLev& L, int p, int i, int j, int k, const NbOff& o, const double* x, double& off, double& diag)
{
    double S[2][2][2], v[27];
    const double* sg = L.sigma + ((long long)(k - 1) * L.n[1] + (j - 1)) * L.n[0] + (i - 1);
    const long long sy = L.n[0], sz = (long long)L.n[0] * L.n[1];
#pragma unroll
    for (int c = 0; c < 8; ++c) S[c >> 2][(c >> 1) & 1][c & 1] = ld_early_nc(sg + (c >> 2) * sz + ((c >> 1) & 1) * sy + (c & 1));
#pragma unroll
    for (int tt = 0; tt < 27; ++tt) {
        const int nz = (tt % 3 != 1) + ((tt / 3) % 3 != 1) + (tt / 9 != 1);
        if (nz >= 2) v[tt] = ld_early(x + p + (o.D[0][tt % 3] + o.D[1][(tt / 3) % 3] + o.D[2][tt / 9]));
    }
    double acc = 0.0, sum = 0.0;
#pragma unroll
    for (int tt = 26; tt >= 0; --tt) {
        const int di = tt % 3, dj = (tt / 3) % 3, dk = tt / 9;
        const int nz = (di != 1) + (dj != 1) + (dk != 1);
        if (nz < 2) continue;
        double w = 0.0;   // sum of sigma over the cells that touch both nodes
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int az = c >> 2, ay = (c >> 1) & 1, ax = c & 1;
            const bool tx = di == 1 || (di == 0 ? ax == 0 : ax == 1), ty = dj == 1 || (dj == 0 ? ay == 0 : ay == 1), tz = dk == 1 || (dk == 0 ? az == 0 : az == 1);
            if (tx && ty && tz) w += S[az][ay][ax];
        }
        acc += w * v[tt];
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) sum += S[c >> 2][(c >> 1) & 1][c & 1];
    off = L.hinv2_12 * acc;
    diag = -4.0 * L.hinv2_12 * sum;
}
__device__ __forceinline__ unsigned ld_early_u8(const unsigned char* p)
{
    unsigned v;
    asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
// BATCH: 0 plain loads; 1 batched loads on the canonical rows only (20 loads, ~56 registers); 2 batched everywhere (~120 registers);
// 3 SPECULATIVE: an interior node issues its flag AND the 20 neighbour values of the canonical row together and decides afterwards
// (one memory round trip instead of two on the path 98 % of the nodes of the finest level take; the rest reloads for the general row)
// SIG: the level may carry sigma-form rows (flag 2: level 0 with variable sigma); a separate instantiation, because the extra path costs
// the constant-sigma kernel 10 registers and a fifth of its speed
template <int BATCH, bool SIG = false>
__device__ __forceinline__ bool node_row(const EbLev& L, int p, int i, int j, int k, int color, const double* x, double& off, double& diag)
{
    const bool inner = i > 0 && i < L.nn[0] - 1 && j > 0 && j < L.nn[1] - 1 && k > 0 && k < L.nn[2] - 1;
    if (BATCH == 3 && inner) {
        const NbOff o = nb_offsets(L, color);
        const unsigned f = ld_early_u8(L.flag + p);
        const double ce = ld_early_nc(L.canon), cc = ld_early_nc(L.canon + 1), cd = ld_early_nc(L.canon + 2);
        const double* xp = x + p;
        double v[27];
#pragma unroll
        for (int tt = 0; tt < 27; ++tt) {
            const int nz = (tt % 3 != 1) + ((tt / 3) % 3 != 1) + (tt / 9 != 1);
            if (nz >= 2) v[tt] = ld_early(xp + (o.D[0][tt % 3] + o.D[1][(tt / 3) % 3] + o.D[2][tt / 9]));
        }
        if (f == 1) {
            double se = 0.0, sc = 0.0;
#pragma unroll
            for (int tt = 26; tt >= 0; --tt) {
                const int nz = (tt % 3 != 1) + ((tt / 3) % 3 != 1) + (tt / 9 != 1);
                if (nz == 2) se += v[tt];
                if (nz == 3) sc += v[tt];
            }
            diag = cd;
            off = ce * se + cc * sc;
            return true;
        }
        if (SIG && f == 2) { sigma_row(L, p, i, j, k, o, x, off, diag); return true; }
        diag = __ldg(L.st + 13 * L.nnode + p);
        if (diag == 0.0) return false;
        off = offdiag_sum_inner<false>(L, p, o, x);
        return true;
    }
    const unsigned fl = L.flag[p];
    if (SIG && fl == 2 && inner) { sigma_row(L, p, i, j, k, nb_offsets(L, color), x, off, diag); return true; }
    if (fl == 1) {
        diag = L.canon[2];
        off = inner ? offdiag_sum_regular_inner<(BATCH == 1 || BATCH == 2)>(L, p, nb_offsets(L, color), x) : offdiag_sum_regular<(BATCH == 1 || BATCH == 2)>(L, i, j, k, x);
        return true;
    }
    diag = __ldg(L.st + 13 * L.nnode + p);
    if (diag == 0.0) return false;
    off = inner ? offdiag_sum_inner<(BATCH == 2)>(L, p, nb_offsets(L, color), x) : offdiag_sum<(BATCH == 2)>(L, p, i, j, k, x);
    return true;
}

// one colour of a Gauss-Seidel sweep (mlndlap_gscolor_sten).  old == x except on levels where a periodic wrap joins two nodes of
// one colour (odd periodic extent): there old is a snapshot taken before the launch.
template <int BATCH, bool SIG>
__global__ void __launch_bounds__(256) k_eb_gs(const EbLev L, double* x, const double* old, const double* __restrict__ rhs, int color)
{
    // block (64, 4): 64 consecutive i/2 of 4 rows j/2; blockIdx.z = k/2
    const int i2 = blockIdx.x * 64 + threadIdx.x, j2 = blockIdx.y * 4 + threadIdx.y, k2 = blockIdx.z;
    const int i = 2 * i2 + (color & 1), j = 2 * j2 + ((color >> 1) & 1), k = 2 * k2 + (color >> 2);
    if (gridDim.x * gridDim.y * gridDim.z <= 1184u) pdl_trigger();   // a single wave: let the next colour's blocks in right away
    pdl_wait();
    if (i >= L.nn[0] || j >= L.nn[1] || k >= L.nn[2]) return;
    const int p = color * L.CS + (k2 * L.H[1] + j2) * L.H[0] + i2;
    const double r = BATCH >= 2 ? ld_early_nc(rhs + p) : rhs[p];   // in flight together with the row's loads
    double off, d;
    x[p] = node_row<BATCH, SIG>(L, p, i, j, k, color, old, off, d) ? (r - off) / d : 0.0;
}
// all sweeps of a smooth call on a level small enough for ONE CTA: colours separated by __syncthreads instead of kernel boundaries
// (a level of a few thousand nodes is pure launch latency otherwise: 8 launches per sweep).  snap != nullptr: odd periodic extent.
__global__ void __launch_bounds__(1024) k_eb_gs_small(const EbLev L, double* x, double* snap, const double* __restrict__ rhs, int nsweeps)
{
    pdl_wait();
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int s = 0; s < nsweeps; ++s)
        for (int c = 0; c < 8; ++c) {
            const double* old = x;
            if (snap) {
                for (int t = tid; t < (int)L.nnode; t += nt) snap[t] = x[t];
                __syncthreads();
                old = snap;
            }
            const int lo = c * L.CS, hi = lo + L.CS;
            for (int p = lo + tid; p < hi; p += nt) {
                int i, j, k;
                if (!ndecode(L, p, i, j, k)) continue;
                if (L.flag[p] == 1) { x[p] = (rhs[p] - offdiag_sum_regular<true>(L, i, j, k, old)) / L.canon[2]; continue; }
                const double d = L.st[13 * L.nnode + p];
                x[p] = d == 0.0 ? 0.0 : (rhs[p] - offdiag_sum<true>(L, p, i, j, k, old)) / d;
            }
            __syncthreads();
        }
}
// The same with the level's vector resident in SHARED memory for the whole launch (loaded once, written back once): between the 8 x
// nsweeps colour phases only __syncthreads, and the 20 / 26 neighbour reads of a node are LDS instead of global loads that the previous
// phase's stores have just invalidated in L1 (a colour phase of k_eb_gs_small is one L2 round trip: 14 - 16 us per sweep on levels of
// 54 and 340 nodes; 8 PDL launches per sweep cost 24 - 25 us on the levels of 2 376 and 17 680 nodes).  Flags, rhs, diagonal and the
// stored rows are read-only and stay in L1.  Same operand order as offdiag_sum<true> / offdiag_sum_regular<true>: same bits.
constexpr int EB_SMEM_MAX_NODES = 28000;   // 224 KB of the 227 KB a CTA may have
__global__ void __launch_bounds__(1024) k_eb_gs_smem(const EbLev L, double* x, const double* __restrict__ rhs, int nsweeps)
{
    extern __shared__ __align__(16) double eb_xs[];
    const int tid = threadIdx.x, nt = blockDim.x, nn = (int)L.nnode;
    pdl_wait();
    for (int t = tid; t < nn; t += nt) eb_xs[t] = x[t];
    __syncthreads();
    const double ce = __ldg(L.canon), cc = __ldg(L.canon + 1), cd = __ldg(L.canon + 2);
    for (int s = 0; s < nsweeps; ++s)
        for (int c = 0; c < 8; ++c) {
            const int lo = c * L.CS, hi = lo + L.CS;
            for (int p = lo + tid; p < hi; p += nt) {
                int i, j, k;
                if (!ndecode(L, p, i, j, k)) continue;
                NbIdx q;
                nb_index(L, i, j, k, q);
                const double r = __ldg(rhs + p);
                if (L.flag[p] == 1) {
                    double se = 0.0, sc = 0.0;
#pragma unroll
                    for (int tt = 26; tt >= 0; --tt) {
                        const int nz = (tt % 3 != 1) + ((tt / 3) % 3 != 1) + (tt / 9 != 1);
                        if (nz == 2) se += eb_xs[q.X[tt % 3] + q.Y[(tt / 3) % 3] + q.Z[tt / 9]];
                        if (nz == 3) sc += eb_xs[q.X[tt % 3] + q.Y[(tt / 3) % 3] + q.Z[tt / 9]];
                    }
                    eb_xs[p] = (r - (ce * se + cc * sc)) / cd;
                    continue;
                }
                const double d = __ldg(L.st + 13 * L.nnode + p);
                if (d == 0.0) { eb_xs[p] = 0.0; continue; }
                double cf[27];
#pragma unroll
                for (int tt = 0; tt < 27; ++tt)
                    if (tt != 13) cf[tt] = ld_early_nc(L.st + (long long)tt * L.nnode + p);
                double ax = 0.0;
#pragma unroll
                for (int tt = 26; tt >= 0; --tt)
                    if (tt != 13) ax += cf[tt] * eb_xs[q.X[tt % 3] + q.Y[(tt / 3) % 3] + q.Z[tt / 9]];
                eb_xs[p] = (r - ax) / d;
            }
            __syncthreads();
        }
    for (int t = tid; t < nn; t += nt) x[t] = eb_xs[t];
}
// out = rhs - A x on the active nodes, 0 elsewhere; optional inf-norm partials
template <int BATCH, bool SIG>
__global__ void __launch_bounds__(256) k_eb_residual(const EbLev L, const double* __restrict__ x, const double* __restrict__ rhs, double* __restrict__ out,
                                                     double* __restrict__ norm_partial)
{
    // block (64, 4) as in k_eb_gs; blockIdx.z = 8 (k/2) + colour: the 8 colours of a pair of planes run next to each other in time, so
    // the phi lines they all read come from L2 once (colour-major order, blockIdx.z = colour * H[2] + k/2, read 1.5 x more DRAM: ncu).
    // norm_partial: one entry per block
    __shared__ double sh[34];
    const int color = blockIdx.z & 7, k2 = blockIdx.z >> 3;
    const int i2 = blockIdx.x * 64 + threadIdx.x, j2 = blockIdx.y * 4 + threadIdx.y;
    const int i = 2 * i2 + (color & 1), j = 2 * j2 + ((color >> 1) & 1), k = 2 * k2 + (color >> 2);
    double r = 0.0;
    pdl_wait();
    if (i < L.nn[0] && j < L.nn[1] && k < L.nn[2]) {
        const int p = color * L.CS + (k2 * L.H[1] + j2) * L.H[0] + i2;
        const double b = BATCH >= 2 ? ld_early_nc(rhs + p) : rhs[p], xc = BATCH >= 2 ? ld_early(x + p) : x[p];
        double off, d;
        if (node_row<BATCH, SIG>(L, p, i, j, k, color, x, off, d)) r = b - (d * xc + off);
        if (out) out[p] = r;
    }
    if (norm_partial) {
        const double amax = eb_block_reduce<true>(fabs(r), sh);
        if (threadIdx.x == 0 && threadIdx.y == 0) norm_partial[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = amax;
    }
}
// crse = (1/8) sum_a w(a) fine(2I + a): full weighting = P^T / 8; 0 on inactive coarse nodes
__global__ void __launch_bounds__(256) k_eb_restrict(const EbLev C, const EbLev F, const double* __restrict__ fine, double* __restrict__ crse)
{
    pdl_wait();
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < C.nnode; t += (long long)gridDim.x * blockDim.x) {
        double s = 0.0;
        if (C.st[13 * C.nnode + t] != 0.0) {
            int I, J, K;
            ndecode(C, t, I, J, K);
            Nb q;
            nb_coords(F, 2 * I, 2 * J, 2 * K, q);
#pragma unroll
            for (int dk = 0; dk < 3; ++dk)
#pragma unroll
                for (int dj = 0; dj < 3; ++dj)
#pragma unroll
                    for (int di = 0; di < 3; ++di) {
                        if (!(q.ok[0][di] && q.ok[1][dj] && q.ok[2][dk])) continue;
                        s += w1(di - 1) * w1(dj - 1) * w1(dk - 1) * fine[nidx(F, q.c[0][di], q.c[1][dj], q.c[2][dk])];
                    }
            s *= 0.125;
        }
        crse[t] = s;
    }
}
// fine += P crse (trilinear) on the active fine nodes
__global__ void __launch_bounds__(256) k_eb_interp_add(const EbLev F, const EbLev C, double* __restrict__ fine, const double* __restrict__ crse)
{
    pdl_wait();
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < F.nnode; t += (long long)gridDim.x * blockDim.x) {
        if (F.st[13 * F.nnode + t] == 0.0) continue;
        int i, j, k;
        ndecode(F, t, i, j, k);
        const int p[3] = {i, j, k};
        int lo[3], hi[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            lo[d] = p[d] >> 1;
            hi[d] = (p[d] + 1) >> 1;
            if (hi[d] >= C.nn[d]) hi[d] = C.per[d] ? 0 : lo[d];
        }
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < 8; ++c)
            s += crse[nidx(C, (c & 1) ? hi[0] : lo[0], (c & 2) ? hi[1] : lo[1], (c & 4) ? hi[2] : lo[2])];
        fine[t] += 0.125 * s;
    }
}

// sums over the active nodes (solvability offset): partial[b] = sum, partial[nb + b] = count
__global__ void __launch_bounds__(256) k_eb_sum_active(const EbLev L, const double* __restrict__ x, double* __restrict__ partial)
{
    __shared__ double sh[34];
    double s = 0.0, c = 0.0;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < L.nnode; p += (long long)gridDim.x * blockDim.x)
        if (L.st[13 * L.nnode + p] != 0.0) { s += x[p]; c += 1.0; }
    s = eb_block_reduce<false>(s, sh);
    c = eb_block_reduce<false>(c, sh);
    if (threadIdx.x == 0) { partial[blockIdx.x] = s; partial[gridDim.x + blockIdx.x] = c; }
}
__global__ void __launch_bounds__(1024) k_eb_mean_final(const double* __restrict__ partial, int nb_, double* __restrict__ out)
{
    __shared__ double sh[34];
    double s = 0.0, c = 0.0;
    for (int t = threadIdx.x; t < nb_; t += blockDim.x) { s += partial[t]; c += partial[nb_ + t]; }
    s = eb_block_reduce<false>(s, sh);
    c = eb_block_reduce<false>(c, sh);
    if (threadIdx.x == 0) out[0] = c > 0.0 ? s / c : 0.0;
}
// x = active ? x - mean : 0, optional inf-norm partials of the result
__global__ void __launch_bounds__(256) k_eb_sub_active(const EbLev L, double* __restrict__ x, const double* __restrict__ mean, double* __restrict__ norm_partial)
{
    __shared__ double sh[34];
    const double m = mean ? mean[0] : 0.0;
    double amax = 0.0;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < L.nnode; p += (long long)gridDim.x * blockDim.x) {
        const double v = L.st[13 * L.nnode + p] != 0.0 ? x[p] - m : 0.0;
        x[p] = v;
        amax = fmax(amax, fabs(v));
    }
    if (norm_partial) {
        amax = eb_block_reduce<true>(amax, sh);
        if (threadIdx.x == 0) norm_partial[blockIdx.x] = amax;
    }
}
__global__ void __launch_bounds__(1024) k_eb_max_final(const double* __restrict__ partial, int nb_, double* __restrict__ out)
{
    __shared__ double sh[34];
    double a = 0.0;
    for (int t = threadIdx.x; t < nb_; t += blockDim.x) a = fmax(a, partial[t]);
    a = eb_block_reduce<true>(a, sh);
    if (threadIdx.x == 0) out[0] = a;
}
__global__ void __launch_bounds__(256) k_eb_axpy(double* __restrict__ y, const double* __restrict__ x, long long n)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) y[t] += x[t];
}

// Bottom solve in ONE CTA: MLCGSolver::solve_bicgstab (x0 = 0, plain dot products, reductions in a fixed order), solvability offset
// over the active nodes first when the operator is singular; on failure 8 smooth calls (MLMG::bottomSolve).
// work: 7 vectors of N doubles.  info[0] += iterations, info[1] = return code.
__global__ void __launch_bounds__(1024) k_eb_bottom(const EbLev L, double* __restrict__ x, double* __restrict__ b, double* __restrict__ work, int maxiter,
                                                    double eps_rel, double eps_abs, int singular, int nsweeps, int* __restrict__ info)
{
    __shared__ double sh[34];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int N = (int)L.nnode;
    double *r = work, *rh = work + N, *pp = work + 2 * (long long)N, *v = work + 3 * (long long)N, *s = work + 4 * (long long)N, *tt = work + 5 * (long long)N,
           *snap = work + 6 * (long long)N;
    const double* dg = L.st + 13 * L.nnode;
    auto apply = [&](const double* in, double* out) {
        __syncthreads();
        for (int t = tid; t < N; t += nt) {
            double y = 0.0;
            if (dg[t] != 0.0) { int i, j, k; ndecode(L, t, i, j, k); y = dg[t] * in[t] + offdiag_sum<true>(L, t, i, j, k, in); }
            out[t] = y;
        }
        __syncthreads();
    };
    auto dot = [&](const double* a, const double* c) { double q = 0.0; for (int t = tid; t < N; t += nt) q += a[t] * c[t]; return eb_block_reduce<false>(q, sh); };
    auto ninf = [&](const double* a) { double q = 0.0; for (int t = tid; t < N; t += nt) q = fmax(q, fabs(a[t])); return eb_block_reduce<true>(q, sh); };
    {
        double q = 0.0, c = 0.0;
        if (singular) {
            for (int t = tid; t < N; t += nt) if (dg[t] != 0.0) { q += b[t]; c += 1.0; }
            q = eb_block_reduce<false>(q, sh);
            c = eb_block_reduce<false>(c, sh);
        }
        const double mean = c > 0.0 ? q / c : 0.0;
        for (int t = tid; t < N; t += nt) b[t] = dg[t] != 0.0 ? b[t] - mean : 0.0;
        __syncthreads();
    }
    for (int t = tid; t < N; t += nt) { r[t] = b[t]; rh[t] = b[t]; x[t] = 0.0; }
    __syncthreads();
    double rnorm = ninf(r);
    const double rnorm0 = rnorm;
    int ret = 0, it = 0;
    if (!(rnorm0 == 0.0 || rnorm0 < eps_abs)) {
        double rho_1 = 0.0, alpha = 0.0, omega = 0.0;
        for (it = 1; it <= maxiter; ++it) {
            const double rho = dot(rh, r);
            if (rho == 0.0) { ret = 1; break; }
            if (it == 1) { for (int t = tid; t < N; t += nt) pp[t] = r[t]; }
            else {
                const double beta = (rho / rho_1) * (alpha / omega);
                for (int t = tid; t < N; t += nt) pp[t] = r[t] + beta * (pp[t] - omega * v[t]);
            }
            apply(pp, v);
            const double rhTv = dot(rh, v);
            if (rhTv == 0.0) { ret = 2; break; }
            alpha = rho / rhTv;
            for (int t = tid; t < N; t += nt) { x[t] += alpha * pp[t]; s[t] = r[t] - alpha * v[t]; }
            __syncthreads();
            rnorm = ninf(s);
            if (rnorm < eps_rel * rnorm0 || rnorm < eps_abs) break;
            apply(s, tt);
            const double t2 = dot(tt, tt);
            if (t2 == 0.0) { ret = 3; break; }
            omega = dot(tt, s) / t2;
            for (int t = tid; t < N; t += nt) { x[t] += omega * s[t]; r[t] = s[t] - omega * tt[t]; }
            __syncthreads();
            rnorm = ninf(r);
            if (rnorm < eps_rel * rnorm0 || rnorm < eps_abs) break;
            if (omega == 0.0) { ret = 4; break; }
            rho_1 = rho;
        }
        if (ret == 0 && !(rnorm < eps_rel * rnorm0 || rnorm < eps_abs)) ret = 8;
        if (it > maxiter) it = maxiter;
    }
    __syncthreads();
    if (ret != 0) {   // start over with 8 smooth calls; every colour reads a snapshot (one CTA: cheap and always safe)
        for (int t = tid; t < N; t += nt) x[t] = 0.0;
        __syncthreads();
        for (int call = 0; call < 8 * nsweeps; ++call)
            for (int c = 0; c < 8; ++c) {
                for (int t = tid; t < N; t += nt) snap[t] = x[t];
                __syncthreads();
                const int lo = c * L.CS, hi = lo + L.CS;
                for (int t = lo + tid; t < hi; t += nt) {
                    if (dg[t] == 0.0) { x[t] = 0.0; continue; }
                    int i, j, k;
                    ndecode(L, t, i, j, k);
                    x[t] = (b[t] - offdiag_sum<true>(L, t, i, j, k, snap)) / dg[t];
                }
                __syncthreads();
            }
    }
    if (tid == 0) { atomicAdd(info, it); info[1] = ret; }
}

// ---- right-hand side, update, copies ------------------------------------------------------------------------------------------
// rhs = D u: natural finite-element rows.  Cells beyond ONE non-periodic face contribute the normal velocity of the first ghost layer
// (0 at walls, the inflow value at inflow faces) with the geometry of the adjacent interior cell; cells beyond two or three faces
// nothing.  ebf != nullptr: + EB inflow, dxinv * (u_eb . n) * int_{EB face} N_a dA.
__global__ void __launch_bounds__(128) k_eb_divu(const EbLev L, const double* __restrict__ geo, EFab vel, const double* __restrict__ ebf, double dxi, double dyi,
                                                 double dzi, double* __restrict__ rhs)
{
    const long long ncell = (long long)L.n[0] * L.n[1] * L.n[2];
    const double dxinv[3] = {dxi, dyi, dzi};
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < L.nnode; t += (long long)gridDim.x * blockDim.x) {
        double r = 0.0;
        if (L.flag[t]) {   // 8 uncut cells, none of them a ghost cell: int_F d_d N_a = s_d / 4
            int i, j, k;
            ndecode(L, t, i, j, k);
            const int p[3] = {i, j, k};
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                int g[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) { g[d] = p[d] - ((a >> d) & 1); if (g[d] < 0) g[d] = L.n[d] - 1; }
                const long long qv = vel.idx(g[0], g[1], g[2]);
#pragma unroll
                for (int d = 0; d < 3; ++d) r -= dxinv[d] * vel.p[d * vel.cs + qv] * (((a >> d) & 1) ? 0.25 : -0.25);
            }
        } else if (L.st[13 * L.nnode + t] != 0.0) {
            int i, j, k;
            ndecode(L, t, i, j, k);
            const int p[3] = {i, j, k};
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                int c[3], g[3];      // cell used for the geometry, cell used for the velocity
                int nout = 0, dout = 0;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    c[d] = g[d] = p[d] - ((a >> d) & 1);
                    if (c[d] < 0) { if (L.per[d]) c[d] = g[d] = L.n[d] - 1; else { c[d] = 0; ++nout; dout = d; } }
                    else if (c[d] >= L.n[d]) { c[d] = L.n[d] - 1; ++nout; dout = d; }   // non-periodic top node only
                }
                if (nout >= 2) continue;
                const long long cc = ((long long)c[2] * L.n[1] + c[1]) * L.n[0] + c[0];
                if (geo[cc] == 0.0) continue;
                double M[3][3][3];
                load_moments(geo, ncell, cc, M);
                const long long qv = vel.idx(g[0], g[1], g[2]);
                if (nout == 1) {
                    r -= dxinv[dout] * vel.p[dout * vel.cs + qv] * grad_integral(M, dout, a);
                } else {
#pragma unroll
                    for (int d = 0; d < 3; ++d) r -= dxinv[d] * vel.p[d * vel.cs + qv] * grad_integral(M, d, a);
                    if (ebf) {
                        const double vn = ebf[cc];
                        if (vn != 0.0) {
                            const double sx = (a & 1) ? 1.0 : -1.0, sy = (a & 2) ? 1.0 : -1.0, sz = (a & 4) ? 1.0 : -1.0;
                            const double* B = ebf + ncell + cc;
                            const double bn = 0.125 * B[0] + 0.25 * (sx * B[ncell] + sy * B[2 * ncell] + sz * B[3 * ncell]) +
                                              0.5 * (sx * sy * B[4 * ncell] + sx * sz * B[5 * ncell] + sy * sz * B[6 * ncell]) + sx * sy * sz * B[7 * ncell];
                            r += dxinv[0] * vn * bn;
                        }
                    }
                }
            }
        }
        rhs[t] = r;
    }
}
// u -= sigma * (1/V) int_F grad phi; gphi = that average (optionally accumulated); covered cells: u = 0, gphi = 0.
// velo: add velocity_o back afterwards (ApplyNodalProjection :85-91).
__global__ void __launch_bounds__(256) k_eb_mknewu(const EbLev L, const double* __restrict__ geo, const double* __restrict__ sigma, const double* __restrict__ phi,
                                                   EFab vel, EFab velo, EFab gphi, int acc_g, double dxi, double dyi, double dzi)
{
    const long long ncell = (long long)L.n[0] * L.n[1] * L.n[2];
    const double dxinv[3] = {dxi, dyi, dzi};
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < ncell; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % L.n[0]), j = (int)((t / L.n[0]) % L.n[1]), k = (int)(t / ((long long)L.n[0] * L.n[1]));
        const double V = geo[t];
        double g[3] = {0.0, 0.0, 0.0};
        if (V != 0.0) {
            double M[3][3][3];
            if (V != 1.0) load_moments(geo, ncell, t, M);   // an uncut cell has the integrals of the cube: int_F d_d N_a = s_d / 4
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                int q[3] = {i + (a & 1), j + ((a >> 1) & 1), k + (a >> 2)};
#pragma unroll
                for (int d = 0; d < 3; ++d) if (q[d] >= L.nn[d]) q[d] = 0;   // periodic wrap
                const double pa = phi[nidx(L, q[0], q[1], q[2])];
#pragma unroll
                for (int d = 0; d < 3; ++d) g[d] += pa * (V != 1.0 ? grad_integral(M, d, a) : (((a >> d) & 1) ? 0.25 : -0.25));
            }
            const double vinv = 1.0 / V;
#pragma unroll
            for (int d = 0; d < 3; ++d) g[d] *= dxinv[d] * vinv;
        }
        const long long qv = vel.idx(i, j, k);
        const double sg = sigma[t];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            double u = V != 0.0 ? vel.p[d * vel.cs + qv] - sg * g[d] : 0.0;
            if (velo.p) u += velo.p[d * velo.cs + velo.idx(i, j, k)];
            vel.p[d * vel.cs + qv] = u;
        }
        if (gphi.p) {
            const long long qg = gphi.idx(i, j, k);
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                if (acc_g) gphi.p[d * gphi.cs + qg] += g[d]; else gphi.p[d * gphi.cs + qg] = g[d];
            }
        }
    }
}
// colour-major node array <-> the caller's nodal fab (natural order): to_fab = 1: fab (=|+=) x; 0: x = fab
__global__ void __launch_bounds__(256) k_eb_copy_nodes(const EbLev L, double* __restrict__ x, EFab f, int to_fab, int accumulate)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < L.nnode; t += (long long)gridDim.x * blockDim.x) {
        int i, j, k;
        if (!ndecode(L, t, i, j, k)) continue;
        const long long q = f.idx(i, j, k);
        if (!to_fab) x[t] = f.p[q];
        else if (accumulate) f.p[q] += x[t];
        else f.p[q] = x[t];
    }
}
// a periodic nodal fab of the caller also holds the duplicate node n (= node 0): fill it after the copy-out
__global__ void __launch_bounds__(256) k_eb_fill_dup(const EbLev L, const double* __restrict__ x, EFab f, int hx, int hy, int hz, int accumulate)
{
    const long long N = (long long)(hx + 1) * (hy + 1) * (hz + 1);
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < N; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % (hx + 1)), j = (int)((t / (hx + 1)) % (hy + 1)), k = (int)(t / ((long long)(hx + 1) * (hy + 1)));
        if (i < L.nn[0] && j < L.nn[1] && k < L.nn[2]) continue;
        const double v = x[nidx(L, i % L.nn[0], j % L.nn[1], k % L.nn[2])];
        const long long q = f.idx(i, j, k);
        if (accumulate) f.p[q] += v; else f.p[q] = v;
    }
}
// the 13 forward entries + the diagonal of a level in natural node order (test hook)
__global__ void __launch_bounds__(256) k_eb_export_stencil(const EbLev L, double* __restrict__ out)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < L.nnode; t += (long long)gridDim.x * blockDim.x) {
        int i, j, k;
        if (!ndecode(L, t, i, j, k)) continue;
        const long long q = ((long long)k * L.nn[1] + j) * L.nn[0] + i;
        const long long nr = (long long)L.nn[0] * L.nn[1] * L.nn[2];
        for (int m = 0; m < 13; ++m) out[(long long)m * nr + q] = L.st[(long long)(14 + m) * L.nnode + t];
        out[13 * nr + q] = L.st[13 * L.nnode + t];
    }
}

// ---- incflo-level kernels (ApplyNodalProjection around the projector) --------------------------------------------------------
// :53-59 u += s * gp / rho (unless incremental); :68 u -= u_old (incremental || proj_for_small_dt); sigma = s / rho (:115-118)
__global__ void __launch_bounds__(256) k_eb_pre_add_sigma(int nx, int ny, int nz, EFab vel, EFab gp, EFab rho, EFab velo, double s, double ro_0, int add_gp,
                                                          int sub_old, double* __restrict__ sigma)
{
    const long long N = (long long)nx * ny * nz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < N; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % nx), j = (int)((t / nx) % ny), k = (int)(t / ((long long)nx * ny));
        const double sor = s / (rho.p ? rho.p[rho.idx(i, j, k)] : ro_0);
        sigma[t] = sor;
        const long long qv = vel.idx(i, j, k);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            double u = vel.p[d * vel.cs + qv];
            if (add_gp) u += gp.p[d * gp.cs + gp.idx(i, j, k)] * sor;
            if (sub_old) u -= velo.p[d * velo.cs + velo.idx(i, j, k)];
            vel.p[d * vel.cs + qv] = u;
        }
    }
}
// vel.setBndry(0) (:137) over every ghost cell of the caller's box, then the first ghost layer of the INFLOW faces from `inflow`
// (same box as vel; :138-163)
__global__ void __launch_bounds__(256) k_eb_set_vel_ghosts(int nx, int ny, int nz, int bx0, int bx1, int by0, int by1, int bz0, int bz1, EFab vel, EFab inflow,
                                                           int inflo0, int inflo1, int inflo2, int infhi0, int infhi1, int infhi2)
{
    const int ex = bx1 - bx0 + 1, ey = by1 - by0 + 1, ez = bz1 - bz0 + 1;
    const long long N = (long long)ex * ey * ez;
    const int n[3] = {nx, ny, nz};
    const int inflo[3] = {inflo0, inflo1, inflo2}, infhi[3] = {infhi0, infhi1, infhi2};
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < N; t += (long long)gridDim.x * blockDim.x) {
        const int q[3] = {bx0 + (int)(t % ex), by0 + (int)((t / ex) % ey), bz0 + (int)(t / ((long long)ex * ey))};
        int nout = 0, dout = 0, side = 0;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if (q[d] < 0) { ++nout; dout = d; side = 0; }
            if (q[d] >= n[d]) { ++nout; dout = d; side = 1; }
        }
        if (nout == 0) continue;
        const long long qv = vel.idx(q[0], q[1], q[2]);
        const bool first = nout == 1 && (side == 0 ? q[dout] == -1 : q[dout] == n[dout]);
        const bool fill = first && inflow.p && (side == 0 ? inflo[dout] : infhi[dout]);
#pragma unroll
        for (int d = 0; d < 3; ++d) vel.p[d * vel.cs + qv] = fill ? inflow.p[d * inflow.cs + qv] : 0.0;
    }
}
// set_eb_velocity / set_eb_density / set_eb_tracer (src/boundary_conditions/incflo_set_bcs.cpp:195-431): 0 everywhere; in cut cells
// (EBCellFlag::isSingleValued) the eb_flow value, masked by the direction test against eb_flow.normal; then FillBoundary of nghost
// layers across periodic faces.  One launch per output array: out has ncomp components; mode 0: velocity (from magnitude: -n * mag,
// "the EB normal points out of the domain"), 1: constant components val[0..ncomp)
__global__ void __launch_bounds__(256) k_eb_set_flow(int nx, int ny, int nz, int per0, int per1, int per2, int nghost, const double* __restrict__ geo, EFab bn,
                                                     EFab out, int bx0, int bx1, int by0, int by1, int bz0, int bz1, int ncomp, int mode, int has_normal,
                                                     double n0, double n1, double n2, double tol_lo, double tol_hi, double mag, const double* __restrict__ val)
{
    const int ex = bx1 - bx0 + 1, ey = by1 - by0 + 1, ez = bz1 - bz0 + 1;
    const long long N = (long long)ex * ey * ez;
    const int n[3] = {nx, ny, nz}, per[3] = {per0, per1, per2};
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < N; t += (long long)gridDim.x * blockDim.x) {
        const int q[3] = {bx0 + (int)(t % ex), by0 + (int)((t / ex) % ey), bz0 + (int)(t / ((long long)ex * ey))};
        int c[3];
        bool ok = true;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            c[d] = q[d];
            if (q[d] < 0 || q[d] >= n[d]) {
                const int dist = q[d] < 0 ? -q[d] : q[d] - n[d] + 1;
                if (per[d] && dist <= nghost) c[d] = (q[d] % n[d] + n[d]) % n[d]; else ok = false;
            }
        }
        const long long qo = out.idx(q[0], q[1], q[2]);
        double mask = 0.0, nv[3] = {0.0, 0.0, 0.0};
        if (ok) {
            const double V = geo[((long long)c[2] * ny + c[1]) * nx + c[0]];
            if (V > 0.0 && V < 1.0) {
                const long long qn = bn.idx(c[0], c[1], c[2]);
                nv[0] = bn.p[qn]; nv[1] = bn.p[bn.cs + qn]; nv[2] = bn.p[2 * bn.cs + qn];
                mask = 1.0;
                if (has_normal) {
                    const double dp = nv[0] * n0 + nv[1] * n1 + nv[2] * n2;
                    mask = (tol_lo <= dp && dp <= tol_hi) ? 1.0 : 0.0;
                }
            }
        }
        for (int m = 0; m < ncomp; ++m) {
            double v = 0.0;
            if (mask != 0.0) v = mode == 0 ? -mask * nv[m] * mag : mask * val[m];
            out.p[m * out.cs + qo] = v;
        }
    }
}
__global__ void __launch_bounds__(256) k_eb_permute(const EbLev L, const double* __restrict__ in, double* __restrict__ out, int to_natural)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < L.nnode; t += (long long)gridDim.x * blockDim.x) {
        int i, j, k;
        if (!ndecode(L, t, i, j, k)) continue;
        const long long q = ((long long)k * L.nn[1] + j) * L.nn[0] + i;
        if (to_natural) out[q] = in[t]; else out[t] = in[q];
    }
}

// ---- multi-box MultiFabs at the boundary (amr.max_grid_size < domain, e.g. test_3d/benchmark.channel_sphere: 16) ----
struct EbMfFab {
    double* p;
    int lo[3];            // allocated box
    int nx, ny, nz;
    long long cs;
};
struct EDense {           // one box over the level with its extent (has())
    double* p;
    int lo[3], hi[3];
    long long cs;
    __device__ __forceinline__ bool has(int i, int j, int k) const { return i >= lo[0] && i <= hi[0] && j >= lo[1] && j <= hi[1] && k >= lo[2] && k <= hi[2]; }
    __device__ __forceinline__ long long idx(int i, int j, int k) const
    {
        return (i - lo[0]) + (long long)(hi[0] - lo[0] + 1) * ((j - lo[1]) + (long long)(hi[1] - lo[1] + 1) * (k - lo[2]));
    }
};
enum { EBMF_VALID = 0,      // the valid boxes only
       EBMF_VALID_BC = 1 }; // gather: + ghost cells OUTSIDE the domain (the BC ghost layer of the velocity); scatter: every cell of the allocated box --
                            // the dense value where there is one (valid cells, first ghost layer), 0 elsewhere (vel.setBndry(0))
__global__ void __launch_bounds__(256) k_eb_mf_gather(const EbMfFab* __restrict__ tab, int ngrow, int ncomp, EDense dst, int mode, int dom0, int dom1, int dom2)
{
    const EbMfFab f = tab[blockIdx.y];
    const long long total = (long long)f.nx * f.ny * f.nz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int li = (int)(t % f.nx), lj = (int)((t / f.nx) % f.ny), lk = (int)(t / ((long long)f.nx * f.ny));
        const int i = li + f.lo[0], j = lj + f.lo[1], k = lk + f.lo[2];
        bool take = li >= ngrow && li < f.nx - ngrow && lj >= ngrow && lj < f.ny - ngrow && lk >= ngrow && lk < f.nz - ngrow;
        if (!take && mode == EBMF_VALID_BC) take = i < 0 || i >= dom0 || j < 0 || j >= dom1 || k < 0 || k >= dom2;
        if (!take || !dst.has(i, j, k)) continue;
        const long long q = dst.idx(i, j, k);
        for (int c = 0; c < ncomp; ++c) dst.p[c * dst.cs + q] = f.p[c * f.cs + t];
    }
}
__global__ void __launch_bounds__(256) k_eb_mf_scatter(const EbMfFab* __restrict__ tab, int ngrow, int ncomp, EDense src, int mode)
{
    const EbMfFab f = tab[blockIdx.y];
    const long long total = (long long)f.nx * f.ny * f.nz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int li = (int)(t % f.nx), lj = (int)((t / f.nx) % f.ny), lk = (int)(t / ((long long)f.nx * f.ny));
        const int i = li + f.lo[0], j = lj + f.lo[1], k = lk + f.lo[2];
        const bool valid = li >= ngrow && li < f.nx - ngrow && lj >= ngrow && lj < f.ny - ngrow && lk >= ngrow && lk < f.nz - ngrow;
        if (!valid && mode != EBMF_VALID_BC) continue;
        const bool have = src.has(i, j, k);
        if (!have && valid) continue;
        const long long q = have ? src.idx(i, j, k) : 0;
        for (int c = 0; c < ncomp; ++c) f.p[c * f.cs + t] = have ? src.p[c * src.cs + q] : 0.0;
    }
}

struct EbLevel {
    EbLev g{};
    unsigned char* flag = nullptr;
    double *cor = nullptr, *res = nullptr, *rescor = nullptr, *sol = nullptr, *rhs = nullptr;
    bool odd_periodic = false;
};

bool eb_is_dev_ptr(const void* p)
{
    if (!p) return true;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
size_t eb_nreal(const EbLev& g) { return (size_t)g.nn[0] * g.nn[1] * g.nn[2]; }   // nodes without the holes: length of a natural-order array
size_t eb_box_doubles(const b200np_fab* b)
{
    return (size_t)(b->hi[0] - b->lo[0] + 1) * (b->hi[1] - b->lo[1] + 1) * (b->hi[2] - b->lo[2] + 1) * (size_t)std::max(b->ncomp, 1);
}
EFab efab(double* p, const b200np_fab* b)
{
    EFab f{};
    f.p = p;
    if (!b || !p) { f.p = nullptr; return f; }
    for (int d = 0; d < 3; ++d) f.lo[d] = b->lo[d];
    f.nx = b->hi[0] - b->lo[0] + 1; f.ny = b->hi[1] - b->lo[1] + 1;
    f.cs = (long long)f.nx * f.ny * (b->hi[2] - b->lo[2] + 1);
    return f;
}
int eb_grid(long long n, int per_block = 256) { return (int)std::max<long long>(1, std::min<long long>((n + per_block - 1) / per_block, 148 * 16)); }
// the box must hold [lo, hi] per direction
bool eb_box_covers(const b200np_fab* b, const int lo[3], const int hi[3], int ncomp)
{
    if (!b || b->ncomp < ncomp) return false;
    for (int d = 0; d < 3; ++d) if (b->lo[d] > lo[d] || b->hi[d] < hi[d]) return false;
    return true;
}

}  // namespace

struct b200eb {
    b200np_geom geom{};
    b200np_opts opts{};
    int device = 0;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    std::vector<EbLevel> lv;
    std::vector<void*> allocs;
    double *geo = nullptr, *ebf = nullptr, *sigma = nullptr;
    double *partial = nullptr, *dscal = nullptr, *work = nullptr, *snap = nullptr, *tmp_nat = nullptr;
    int* dinfo = nullptr;
    double* hscal = nullptr;
    int* hinfo = nullptr;
    bool singular = true, have_geometry = false, have_ebflow = false, have_stencil = false;
    int flags_state = 0;      // 0: unknown; see eb_build_stencils
    double* canon = nullptr;  // 3 doubles per level
    long long batch_below = 500000;    // levels with fewer nodes use the load-batching kernels (B200EB_BATCH_BELOW)
    int use_pdl = 1;          // programmatic dependent launch between the V-cycle kernels (B200EB_PDL)
    int big_variant = 3;      // levels of >= batch_below nodes: 0 plain loads, 1 batched loads on the canonical rows, 3 speculative (B200EB_BIG_VARIANT)
    int small_nodes = 1000;   // levels up to this many nodes smooth in one CTA (B200EB_SMALL_NODES)
    int smem_nodes = 3000;    // ... with the vector resident in shared memory (k_eb_gs_smem; B200EB_SMEM_NODES, at most EB_SMEM_MAX_NODES)
    long long launches = 0, ncell = 0;
    struct Stage { double* d = nullptr; size_t bytes = 0; } stage[12];
    struct Mf { double* dense = nullptr; size_t dense_doubles = 0; double* stage = nullptr; size_t stage_doubles = 0; EbMfFab* tab = nullptr; size_t tab_cap = 0; } mf[10];
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    long long launches_per_vcycle = 0;
};

namespace {

#define ELAUNCH(h, kern, grid, block, ...)                  \
    do {                                                    \
        kern<<<grid, block, 0, (h)->stream>>>(__VA_ARGS__); \
        (h)->launches++;                                    \
    } while (0)

dim3 eb_grid3(const EbLev& g, int ncolors) { return dim3((g.H[0] + 63) / 64, (g.H[1] + 3) / 4, g.H[2] * ncolors); }
int eb_blocks3(const EbLev& g, int ncolors) { const dim3 d = eb_grid3(g, ncolors); return (int)(d.x * d.y * d.z); }

// launch with the programmatic-stream-serialization attribute: ONLY for kernels that call pdl_wait() before touching data
template <typename... KArgs, typename... Args>
void eb_launch_pdl_smem(b200eb* h, size_t smem, void (*kern)(KArgs...), dim3 grid, dim3 block, Args... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = h->use_pdl ? 1 : 0;
    ECK(cudaLaunchKernelEx(&cfg, kern, KArgs(args)...));
    h->launches++;
}
template <typename... KArgs, typename... Args>
void eb_launch_pdl(b200eb* h, void (*kern)(KArgs...), dim3 grid, dim3 block, Args... args)
{
    eb_launch_pdl_smem(h, 0, kern, grid, block, args...);
}

double* eb_alloc(b200eb* h, size_t doubles)
{
    void* p = nullptr;
    ECK(cudaMalloc(&p, std::max<size_t>(doubles, 1) * sizeof(double)));
    ECK(cudaMemsetAsync(p, 0, std::max<size_t>(doubles, 1) * sizeof(double), h->stream));
    h->allocs.push_back(p);
    return static_cast<double*>(p);
}

void eb_build(b200eb* h)
{
    const b200np_geom& G = h->geom;
    int n[3] = {G.n_cell[0], G.n_cell[1], G.n_cell[2]};
    h->singular = true;
    for (int d = 0; d < 3; ++d)
        if (G.bc_lo[d] == B200NP_BC_DIRICHLET || G.bc_hi[d] == B200NP_BC_DIRICHLET) h->singular = false;
    h->ncell = (long long)n[0] * n[1] * n[2];
    for (int lev = 0;; ++lev) {
        EbLevel L;
        EbLev& g = L.g;
        for (int d = 0; d < 3; ++d) {
            g.n[d] = n[d];
            g.per[d] = G.bc_lo[d] == B200NP_BC_PERIODIC;
            g.nn[d] = g.per[d] ? n[d] : n[d] + 1;
            g.dirlo[d] = G.bc_lo[d] == B200NP_BC_DIRICHLET;
            g.dirhi[d] = G.bc_hi[d] == B200NP_BC_DIRICHLET;
            g.H[d] = (g.nn[d] + 1) / 2;
            if (g.per[d] && (g.nn[d] & 1)) L.odd_periodic = true;
        }
        if ((long long)g.H[0] * g.H[1] * g.H[2] * 8 >= (1LL << 31)) throw int(B200NP_ERR_UNSUPPORTED);   // 32-bit node positions
        g.CS = g.H[0] * g.H[1] * g.H[2];
        g.nnode = 8LL * g.CS;
        g.st = eb_alloc(h, (size_t)27 * g.nnode);
        L.flag = reinterpret_cast<unsigned char*>(eb_alloc(h, (size_t)(g.nnode + 7) / 8 + 1));
        g.flag = L.flag;
        L.cor = eb_alloc(h, g.nnode); L.res = eb_alloc(h, g.nnode); L.rescor = eb_alloc(h, g.nnode);
        if (lev == 0) { L.sol = eb_alloc(h, g.nnode); L.rhs = eb_alloc(h, g.nnode); }
        h->lv.push_back(L);
        bool can = lev + 1 <= h->opts.mg_max_coarsening_level && lev + 1 < 30;
        for (int d = 0; d < 3; ++d) if (n[d] % 2 != 0 || n[d] / 2 < 2) can = false;
        if (!can) break;
        for (int d = 0; d < 3; ++d) n[d] /= 2;
    }
    h->canon = eb_alloc(h, 3 * h->lv.size());
    for (size_t l = 0; l < h->lv.size(); ++l) h->lv[l].g.canon = h->canon + 3 * l;
    if (const char* e = getenv("B200EB_SMALL_NODES")) h->small_nodes = atoi(e);
    if (const char* e = getenv("B200EB_SMEM_NODES")) h->smem_nodes = atoi(e);
    h->smem_nodes = std::min(h->smem_nodes, EB_SMEM_MAX_NODES);
    ECK(cudaFuncSetAttribute(k_eb_gs_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, EB_SMEM_MAX_NODES * (int)sizeof(double)));
    if (const char* e = getenv("B200EB_BATCH_BELOW")) h->batch_below = atoll(e);
    if (const char* e = getenv("B200EB_PDL")) h->use_pdl = atoi(e);
    if (const char* e = getenv("B200EB_BIG_VARIANT")) h->big_variant = atoi(e);
    h->geo = eb_alloc(h, (size_t)19 * h->ncell);
    h->sigma = eb_alloc(h, (size_t)h->ncell);
    for (size_t l = 0; l < h->lv.size(); ++l) {
        const double hl = G.dx[0] * (double)(1 << l);
        h->lv[l].g.sigma = l == 0 ? h->sigma : nullptr;
        h->lv[l].g.hinv2_12 = 1.0 / (12.0 * hl * hl);
    }
    h->work = eb_alloc(h, (size_t)7 * h->lv.back().g.nnode);
    h->snap = eb_alloc(h, (size_t)h->lv[0].g.nnode);
    h->tmp_nat = eb_alloc(h, (size_t)14 * h->lv[0].g.nnode);
    h->partial = eb_alloc(h, std::max(2 * 148 * 16, eb_blocks3(h->lv[0].g, 8)) + 8);
    h->dscal = eb_alloc(h, 16);
    h->dinfo = reinterpret_cast<int*>(eb_alloc(h, 4));
    ECK(cudaMallocHost(&h->hscal, 16 * sizeof(double)));
    ECK(cudaMallocHost(&h->hinfo, 8 * sizeof(int)));
}

// MLNodeLaplacian::buildStencil: level 0 from sigma and the integrals, then the Galerkin products.
// const_sigma > 0: sigma is that constant -- nodes away from the body and the domain faces then share one row per level, which is
// flagged instead of stored (B200EB_FLAGS=0 disables it).
void eb_build_stencils(b200eb* h, double const_sigma)
{
    const b200np_geom& G = h->geom;
    EbLevel& L0 = h->lv[0];
    static const bool use_flags = !(getenv("B200EB_FLAGS") && atoi(getenv("B200EB_FLAGS")) == 0);
    // 2: constant sigma -- canonical rows on every level; 1: variable sigma -- sigma-form rows (flag 2) on level 0, stored rows below;
    // 3: no flags at all
    const int want = !use_flags ? 3 : (const_sigma > 0 ? 2 : 1);
    if (h->flags_state != want) {
        if (want == 3) {
            for (auto& L : h->lv) ECK(cudaMemsetAsync(L.flag, 0, (size_t)L.g.nnode, h->stream));
        } else {
            ELAUNCH(h, k_eb_flag0, eb_grid(L0.g.nnode), 256, L0.g, (const double*)h->geo, L0.flag, want == 2 ? 1 : 2);
            for (size_t l = 0; l + 1 < h->lv.size(); ++l) {
                if (want == 2) ELAUNCH(h, k_eb_flag_coarse, eb_grid(h->lv[l + 1].g.nnode), 256, h->lv[l + 1].g, h->lv[l].g, h->lv[l + 1].flag);
                else ECK(cudaMemsetAsync(h->lv[l + 1].flag, 0, (size_t)h->lv[l + 1].g.nnode, h->stream));
            }
        }
        h->flags_state = want;
    }
    if (want == 2) ELAUNCH(h, k_eb_set_canon, 1, 32, h->canon, (int)h->lv.size(), const_sigma, 1.0 / (G.dx[0] * G.dx[0]));
    ELAUNCH(h, k_eb_stencil0, eb_grid(L0.g.nnode, 128), 128, L0.g, (const double*)h->geo, (const double*)h->sigma, 1.0 / (G.dx[0] * G.dx[0]),
            1.0 / (G.dx[1] * G.dx[1]), 1.0 / (G.dx[2] * G.dx[2]));
    for (size_t l = 0; l + 1 < h->lv.size(); ++l)
        ELAUNCH(h, k_eb_rap, eb_grid(h->lv[l + 1].g.nnode, 128), 128, h->lv[l + 1].g, h->lv[l].g);
    h->have_stencil = true;
}

// kernel variant by level size (struct b200eb: batch_below, big_variant)
template <bool SIG>
void eb_launch_gs_t(b200eb* h, EbLevel& L, double* x, const double* old, const double* rhs, int c)
{
    const dim3 grid = eb_grid3(L.g, 1), block(64, 4);
    if (L.g.nnode < h->batch_below) eb_launch_pdl(h, k_eb_gs<2, SIG>, grid, block, L.g, x, old, rhs, c);
    else if (h->big_variant == 3) eb_launch_pdl(h, k_eb_gs<3, SIG>, grid, block, L.g, x, old, rhs, c);
    else if (h->big_variant == 1) eb_launch_pdl(h, k_eb_gs<1, SIG>, grid, block, L.g, x, old, rhs, c);
    else eb_launch_pdl(h, k_eb_gs<0, SIG>, grid, block, L.g, x, old, rhs, c);
}
// sigma-form rows exist on level 0 when sigma is variable (flags_state 1)
void eb_launch_gs(b200eb* h, EbLevel& L, double* x, const double* old, const double* rhs, int c)
{
    if (h->flags_state == 1 && &L == &h->lv[0]) eb_launch_gs_t<true>(h, L, x, old, rhs, c);
    else eb_launch_gs_t<false>(h, L, x, old, rhs, c);
}
template <bool SIG>
void eb_launch_residual_t(b200eb* h, EbLevel& L, const double* x, const double* rhs, double* out, double* partial)
{
    const dim3 grid = eb_grid3(L.g, 8), block(64, 4);
    if (L.g.nnode < h->batch_below) eb_launch_pdl(h, k_eb_residual<2, SIG>, grid, block, L.g, x, rhs, out, partial);
    else if (h->big_variant == 3) eb_launch_pdl(h, k_eb_residual<3, SIG>, grid, block, L.g, x, rhs, out, partial);
    else if (h->big_variant == 1) eb_launch_pdl(h, k_eb_residual<1, SIG>, grid, block, L.g, x, rhs, out, partial);
    else eb_launch_pdl(h, k_eb_residual<0, SIG>, grid, block, L.g, x, rhs, out, partial);
}
void eb_launch_residual(b200eb* h, EbLevel& L, const double* x, const double* rhs, double* out, double* partial)
{
    if (h->flags_state == 1 && &L == &h->lv[0]) eb_launch_residual_t<true>(h, L, x, rhs, out, partial);
    else eb_launch_residual_t<false>(h, L, x, rhs, out, partial);
}

// one MLMG smooth call = smooth_num_sweeps sweeps of 8 colours
void eb_smooth(b200eb* h, EbLevel& L, double* x, const double* rhs, int ncalls)
{
    const int nsw = std::max(1, h->opts.smooth_num_sweeps);
    if (L.g.nnode <= h->smem_nodes && !L.odd_periodic) {   // one CTA, the vector in shared memory
        eb_launch_pdl_smem(h, (size_t)L.g.nnode * sizeof(double), k_eb_gs_smem, dim3(1), dim3(1024), L.g, x, rhs, ncalls * nsw);
        return;
    }
    if (L.g.nnode <= h->small_nodes) {
        eb_launch_pdl(h, k_eb_gs_small, dim3(1), dim3(1024), L.g, x, L.odd_periodic ? h->snap : (double*)nullptr, rhs, ncalls * nsw);
        return;
    }
    for (int s = 0; s < ncalls * nsw; ++s)
        for (int c = 0; c < 8; ++c) {
            const double* old = x;
            if (L.odd_periodic) {
                ECK(cudaMemcpyAsync(h->snap, x, L.g.nnode * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
                old = h->snap;
            }
            eb_launch_gs(h, L, x, old, rhs, c);
        }
}

void eb_bottom(b200eb* h)
{
    EbLevel& B = h->lv.back();
    ELAUNCH(h, k_eb_bottom, 1, 1024, B.g, B.cor, B.res, h->work, h->opts.bottom_maxiter, h->opts.bottom_rtol, h->opts.bottom_atol, h->singular ? 1 : 0,
            std::max(1, h->opts.smooth_num_sweeps), h->dinfo);
}

void eb_vcycle(b200eb* h)
{
    const int nl = (int)h->lv.size();
    const int nu1 = h->opts.num_pre_smooth, nu2 = h->opts.num_post_smooth;
    for (int l = 0; l < nl - 1; ++l) {
        EbLevel &L = h->lv[l], &C = h->lv[l + 1];
        ECK(cudaMemsetAsync(L.cor, 0, L.g.nnode * sizeof(double), h->stream));
        eb_smooth(h, L, L.cor, L.res, nu1);
        eb_launch_residual(h, L, (const double*)L.cor, (const double*)L.res, L.rescor, (double*)nullptr);
        eb_launch_pdl(h, k_eb_restrict, dim3(eb_grid(C.g.nnode)), dim3(256), C.g, L.g, (const double*)L.rescor, C.res);
    }
    eb_bottom(h);
    for (int l = nl - 2; l >= 0; --l) {
        EbLevel &L = h->lv[l], &C = h->lv[l + 1];
        eb_launch_pdl(h, k_eb_interp_add, dim3(eb_grid(L.g.nnode)), dim3(256), L.g, C.g, L.cor, (const double*)C.cor);
        eb_smooth(h, L, L.cor, L.res, nu2);
    }
}

void eb_vcycle_run(b200eb* h)
{
    if (!h->opts.use_graph) { eb_vcycle(h); return; }
    if (!h->graph_exec) {
        const long long before = h->launches;
        ECK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        try { eb_vcycle(h); }
        catch (int) {
            cudaGraph_t broken = nullptr;
            cudaStreamEndCapture(h->stream, &broken);
            if (broken) cudaGraphDestroy(broken);
            cudaGetLastError();
            throw;
        }
        ECK(cudaStreamEndCapture(h->stream, &h->graph));
        ECK(cudaGraphInstantiate(&h->graph_exec, h->graph, 0));
        h->launches_per_vcycle = h->launches - before;
        h->launches = before;
    }
    ECK(cudaGraphLaunch(h->graph_exec, h->stream));
    h->launches += h->launches_per_vcycle;
}

double eb_read_norm(b200eb* h, int nb_)
{
    ELAUNCH(h, k_eb_max_final, 1, 1024, (const double*)h->partial, nb_, h->dscal + 2);
    ECK(cudaMemcpyAsync(h->hscal + 2, h->dscal + 2, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    ECK(cudaStreamSynchronize(h->stream));
    return h->hscal[2];
}

// MLMG::solve on (sol = 0, rhs) of level 0
int eb_solve(b200eb* h, double rtol, double atol, b200np_stats* st)
{
    EbLevel& L0 = h->lv[0];
    st->iters = 0; st->bottom_iters = 0; st->status = B200NP_OK; st->nlevels = (int)h->lv.size();
    ECK(cudaMemsetAsync(h->dinfo, 0, 4 * sizeof(int), h->stream));
    const int nb_ = eb_grid(L0.g.nnode);
    if (h->singular) {   // makeSolvable: remove the mean over the active nodes
        ELAUNCH(h, k_eb_sum_active, nb_, 256, L0.g, (const double*)L0.rhs, h->partial);
        ELAUNCH(h, k_eb_mean_final, 1, 1024, (const double*)h->partial, nb_, h->dscal);
        ELAUNCH(h, k_eb_sub_active, nb_, 256, L0.g, L0.rhs, (const double*)h->dscal, h->partial);
    } else {
        ELAUNCH(h, k_eb_sub_active, nb_, 256, L0.g, L0.rhs, (const double*)nullptr, h->partial);
    }
    st->rhsnorm = eb_read_norm(h, nb_);
    // zero initial guess: the initial residual is rhs itself
    ECK(cudaMemsetAsync(L0.sol, 0, L0.g.nnode * sizeof(double), h->stream));
    ECK(cudaMemcpyAsync(L0.res, L0.rhs, L0.g.nnode * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    st->resnorm0 = st->rhsnorm;
    const double maxnorm = std::max(st->rhsnorm, st->resnorm0);
    const double target = std::max(atol, std::max(rtol, 1e-16) * maxnorm);
    st->resnorm = st->resnorm0;
    st->resnorm_hist[0] = st->resnorm0;
    if (h->opts.verbose >= 1) printf("MLMG: Initial rhs               = %.12g\nMLMG: Initial residual (resid0) = %.12g\n", st->rhsnorm, st->resnorm0);
    if (st->resnorm0 <= target) return B200NP_OK;
    bool converged = false;
    for (int it = 0; it < h->opts.maxiter; ++it) {
        if (h->lv.size() == 1) { eb_bottom(h); }
        else eb_vcycle_run(h);
        ELAUNCH(h, k_eb_axpy, nb_, 256, L0.sol, (const double*)L0.cor, L0.g.nnode);
        eb_launch_residual(h, L0, (const double*)L0.sol, (const double*)L0.rhs, L0.res, h->partial);
        st->resnorm = eb_read_norm(h, eb_blocks3(L0.g, 8));
        st->iters = it + 1;
        if (it + 1 < 128) st->resnorm_hist[it + 1] = st->resnorm;
        if (h->opts.verbose >= 2) printf("MLMG: Iteration %3d Fine resid/bnorm = %.12g\n", it + 1, st->resnorm / maxnorm);
        if (st->resnorm <= target) { converged = true; break; }
        if (!(st->resnorm <= 1e20 * maxnorm)) { st->status = B200NP_ERR_DIVERGED; break; }
    }
    if (!converged && st->status == B200NP_OK) st->status = B200NP_ERR_NOT_CONVERGED;
    ECK(cudaMemcpyAsync(h->hinfo, h->dinfo, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    ECK(cudaStreamSynchronize(h->stream));
    st->bottom_iters = h->hinfo[0];
    if (h->opts.verbose >= 1) printf("MLMG: Final Iter. %d resid, resid/bnorm = %.12g, %.12g\n", st->iters, st->resnorm, st->resnorm / maxnorm);
    return st->status;
}

double* eb_stage_in(b200eb* h, int slot, const double* p, const b200np_fab* box, bool copy, bool* staged, b200np_stats* st)
{
    *staged = false;
    if (!p) return nullptr;
    if (eb_is_dev_ptr(p)) return const_cast<double*>(p);
    const size_t bytes = eb_box_doubles(box) * sizeof(double);
    auto& S = h->stage[slot];
    if (S.bytes < bytes) {
        if (S.d) ECK(cudaFree(S.d));
        ECK(cudaMalloc(&S.d, bytes));
        S.bytes = bytes;
    }
    if (copy) { ECK(cudaMemcpyAsync(S.d, p, bytes, cudaMemcpyHostToDevice, h->stream)); if (st) st->h2d_bytes += (long long)bytes; }
    *staged = true;
    return S.d;
}
void eb_stage_out(b200eb* h, int slot, double* p, const b200np_fab* box, bool staged, b200np_stats* st)
{
    if (!staged || !p) return;
    const size_t bytes = eb_box_doubles(box) * sizeof(double);
    ECK(cudaMemcpyAsync(p, h->stage[slot].d, bytes, cudaMemcpyDeviceToHost, h->stream));
    if (st) st->d2h_bytes += (long long)bytes;
}

// NodalProjector::project on device arrays: sigma already in h->sigma
int eb_project_dev(b200eb* h, double const_sigma, EFab fvel, EFab fvelo, EFab fphi, const b200np_fab* phi_box, int acc_p, EFab fgphi, int acc_g, double rtol,
                   double atol, b200np_stats* st)
{
    const b200np_geom& G = h->geom;
    EbLevel& L0 = h->lv[0];
    const double dxi = 1.0 / G.dx[0], dyi = 1.0 / G.dx[1], dzi = 1.0 / G.dx[2];
    eb_build_stencils(h, const_sigma);
    ELAUNCH(h, k_eb_divu, eb_grid(L0.g.nnode, 128), 128, L0.g, (const double*)h->geo, fvel, (const double*)(h->have_ebflow ? h->ebf : nullptr), dxi, dyi, dzi, L0.rhs);
    ECK(cudaEventRecord(h->ev[2], h->stream));
    const int status = eb_solve(h, rtol, atol, st);
    ECK(cudaEventRecord(h->ev[3], h->stream));
    ELAUNCH(h, k_eb_mknewu, eb_grid(h->ncell), 256, L0.g, (const double*)h->geo, (const double*)h->sigma, (const double*)L0.sol, fvel, fvelo, fgphi, acc_g, dxi, dyi,
            dzi);
    if (fphi.p) {
        ELAUNCH(h, k_eb_copy_nodes, eb_grid(L0.g.nnode), 256, L0.g, L0.sol, fphi, 1, acc_p);
        if (L0.g.per[0] || L0.g.per[1] || L0.g.per[2]) {
            const int hx = std::min(phi_box->hi[0], G.n_cell[0]), hy = std::min(phi_box->hi[1], G.n_cell[1]), hz = std::min(phi_box->hi[2], G.n_cell[2]);
            if (hx >= L0.g.nn[0] || hy >= L0.g.nn[1] || hz >= L0.g.nn[2])
                ELAUNCH(h, k_eb_fill_dup, eb_grid((long long)(hx + 1) * (hy + 1) * (hz + 1)), 256, L0.g, (const double*)L0.sol, fphi, hx, hy, hz, acc_p);
        }
    }
    return status;
}

bool eb_cell_box_ok(const b200eb* h, const b200np_fab* b, int ng, int ncomp)
{
    const int lo[3] = {-ng, -ng, -ng};
    const int hi[3] = {h->geom.n_cell[0] - 1 + ng, h->geom.n_cell[1] - 1 + ng, h->geom.n_cell[2] - 1 + ng};
    return eb_box_covers(b, lo, hi, ncomp);
}
bool eb_node_box_ok(const b200eb* h, const b200np_fab* b)
{
    const int lo[3] = {0, 0, 0};
    const int hi[3] = {h->lv[0].g.nn[0] - 1, h->lv[0].g.nn[1] - 1, h->lv[0].g.nn[2] - 1};
    return eb_box_covers(b, lo, hi, 1);
}

}  // namespace

extern "C" {

int b200eb_create(b200eb_t** out, const b200np_geom* geom, const b200np_opts* opts, int device)
{
    if (!out || !geom) return B200NP_ERR_BAD_ARG;
    *out = nullptr;
    for (int d = 0; d < 3; ++d) {
        if (geom->n_cell[d] < 2 || !(geom->dx[d] > 0)) return B200NP_ERR_BAD_ARG;
        if (geom->bc_lo[d] < 0 || geom->bc_lo[d] > 3 || geom->bc_hi[d] < 0 || geom->bc_hi[d] > 3) return B200NP_ERR_BAD_BC;
        if ((geom->bc_lo[d] == B200NP_BC_PERIODIC) != (geom->bc_hi[d] == B200NP_BC_PERIODIC)) return B200NP_ERR_BAD_BC;
    }
    // AMReX's EB support asserts dx == dy == dz
    if (std::fabs(geom->dx[0] - geom->dx[1]) > 1e-12 * geom->dx[0] || std::fabs(geom->dx[0] - geom->dx[2]) > 1e-12 * geom->dx[0]) return B200NP_ERR_BAD_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) { cudaGetLastError(); return B200NP_ERR_CUDA; }
    b200eb* h = new b200eb();
    try {
        ECK(cudaSetDevice(device));
        h->device = device;
        h->geom = *geom;
        if (opts) h->opts = *opts; else b200np_default_opts(&h->opts);
        ECK(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
        h->stream = h->own_stream;
        for (auto& e : h->ev) ECK(cudaEventCreate(&e));
        eb_build(h);
        ECK(cudaStreamSynchronize(h->stream));
    } catch (int e) { b200eb_destroy(h); return e; }
    *out = h;
    return B200NP_OK;
}

void b200eb_destroy(b200eb_t* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
    if (h->graph) cudaGraphDestroy(h->graph);
    for (void* p : h->allocs) cudaFree(p);
    for (auto& s : h->stage) if (s.d) cudaFree(s.d);
    for (auto& m : h->mf) { if (m.dense) cudaFree(m.dense); if (m.stage) cudaFree(m.stage); if (m.tab) cudaFree(m.tab); }
    if (h->hscal) cudaFreeHost(h->hscal);
    if (h->hinfo) cudaFreeHost(h->hinfo);
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

// run on the caller's CUDA stream (e.g. amrex::Gpu::gpuStream()); NULL restores the handle's own non-blocking stream
int b200eb_set_stream(b200eb_t* h, void* stream)
{
    if (!h) return B200NP_ERR_BAD_ARG;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    h->stream = stream ? (cudaStream_t)stream : h->own_stream;
    return B200NP_OK;
}

int b200eb_nlevels(const b200eb_t* h) { return h ? (int)h->lv.size() : 0; }

int b200eb_set_geometry(b200eb_t* h, const double* vfrac, const b200np_fab* vfrac_box, const double* intg, const b200np_fab* intg_box)
{
    if (!h || !vfrac || !intg) return B200NP_ERR_BAD_ARG;
    if (!eb_cell_box_ok(h, vfrac_box, 0, 1) || !eb_cell_box_ok(h, intg_box, 0, 18)) return B200NP_ERR_BAD_ARG;
    try {
        ECK(cudaSetDevice(h->device));
        bool s0, s1;
        double* dv = eb_stage_in(h, 0, vfrac, vfrac_box, true, &s0, nullptr);
        double* di = eb_stage_in(h, 1, intg, intg_box, true, &s1, nullptr);
        const int* n = h->geom.n_cell;
        ELAUNCH(h, k_eb_copy_geo, eb_grid(h->ncell), 256, n[0], n[1], n[2], efab(dv, vfrac_box), efab(di, intg_box), h->geo);
        ECK(cudaStreamSynchronize(h->stream));
        h->have_geometry = true;
        h->flags_state = 0;
        return B200NP_OK;
    } catch (int e) { return e; }
}

int b200eb_set_eb_inflow_velocity(b200eb_t* h, const double* eb_vel, const b200np_fab* vel_box, const double* bnorm, const b200np_fab* bnorm_box,
                                  const double* bintg, const b200np_fab* bintg_box)
{
    if (!h) return B200NP_ERR_BAD_ARG;
    if (!eb_vel) { h->have_ebflow = false; return B200NP_OK; }
    if (!bnorm || !bintg) return B200NP_ERR_BAD_ARG;
    if (!eb_cell_box_ok(h, vel_box, 0, 3) || !eb_cell_box_ok(h, bnorm_box, 0, 3) || !eb_cell_box_ok(h, bintg_box, 0, 8)) return B200NP_ERR_BAD_ARG;
    try {
        ECK(cudaSetDevice(h->device));
        if (!h->ebf) h->ebf = eb_alloc(h, (size_t)9 * h->ncell);
        bool s0, s1, s2;
        double* dv = eb_stage_in(h, 0, eb_vel, vel_box, true, &s0, nullptr);
        double* dn = eb_stage_in(h, 1, bnorm, bnorm_box, true, &s1, nullptr);
        double* db = eb_stage_in(h, 2, bintg, bintg_box, true, &s2, nullptr);
        const int* n = h->geom.n_cell;
        ELAUNCH(h, k_eb_copy_flow, eb_grid(h->ncell), 256, n[0], n[1], n[2], efab(dv, vel_box), efab(dn, bnorm_box), efab(db, bintg_box), h->ebf);
        ECK(cudaStreamSynchronize(h->stream));
        h->have_ebflow = true;
        return B200NP_OK;
    } catch (int e) { return e; }
}

int b200eb_set_eb_flow(b200eb_t* h, const b200eb_flow* f, int nghost, const double* bnorm, const b200np_fab* bnorm_box, double* eb_vel,
                       const b200np_fab* vel_box, double* eb_density, const b200np_fab* density_box, double* eb_tracer, const b200np_fab* tracer_box)
{
    if (!h || !f || !bnorm || !h->have_geometry || nghost < 0) return B200NP_ERR_BAD_ARG;
    if (!eb_cell_box_ok(h, bnorm_box, 0, 3)) return B200NP_ERR_BAD_ARG;
    if (f->ntrac < 0 || f->ntrac > 8) return B200NP_ERR_BAD_ARG;
    if ((eb_vel && !eb_cell_box_ok(h, vel_box, 0, 3)) || (eb_density && !eb_cell_box_ok(h, density_box, 0, 1)) ||
        (eb_tracer && !eb_cell_box_ok(h, tracer_box, 0, std::max(f->ntrac, 1))))
        return B200NP_ERR_BAD_ARG;
    try {
        ECK(cudaSetDevice(h->device));
        bool sn;
        double* dn = eb_stage_in(h, 0, bnorm, bnorm_box, true, &sn, nullptr);
        const EFab fn = efab(dn, bnorm_box);
        const int* n = h->geom.n_cell;
        const int per[3] = {h->lv[0].g.per[0], h->lv[0].g.per[1], h->lv[0].g.per[2]};
        // Real pad = std::numeric_limits<float>::epsilon(); norm_tol_lo / hi = -1 -/+ (normal_tol + pad)   (:218-221)
        const double pad = 1.1920928955078125e-07;
        const double tol_lo = -1.0 - (f->normal_tol + pad), tol_hi = -1.0 + (f->normal_tol + pad);
        double* dval = h->dscal + 4;   // up to 8 constants
        struct Out { double* p; const b200np_fab* box; int ncomp; int mode; double val[8]; } outs[3] = {
            {eb_vel, vel_box, 3, f->is_mag ? 0 : 1, {f->velocity[0], f->velocity[1], f->velocity[2]}},
            {eb_density, density_box, 1, 1, {f->density}},
            {eb_tracer, tracer_box, f->ntrac, 1, {}}};
        for (int m = 0; m < f->ntrac; ++m) outs[2].val[m] = f->tracer[m];
        for (int o = 0; o < 3; ++o) {
            if (!outs[o].p || outs[o].ncomp == 0) continue;
            bool so;
            double* dp = eb_stage_in(h, 1 + o, outs[o].p, outs[o].box, false, &so, nullptr);
            ECK(cudaMemcpyAsync(dval, outs[o].val, 8 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            ECK(cudaStreamSynchronize(h->stream));   // outs[o].val is a stack array
            const b200np_fab* b = outs[o].box;
            const long long N = (long long)(b->hi[0] - b->lo[0] + 1) * (b->hi[1] - b->lo[1] + 1) * (b->hi[2] - b->lo[2] + 1);
            ELAUNCH(h, k_eb_set_flow, eb_grid(N), 256, n[0], n[1], n[2], per[0], per[1], per[2], nghost, (const double*)h->geo, fn, efab(dp, b), b->lo[0], b->hi[0],
                    b->lo[1], b->hi[1], b->lo[2], b->hi[2], outs[o].ncomp, outs[o].mode, f->has_normal, f->normal[0], f->normal[1], f->normal[2], tol_lo, tol_hi,
                    f->vel_mag, (const double*)dval);
            eb_stage_out(h, 1 + o, outs[o].p, b, so, nullptr);
            ECK(cudaStreamSynchronize(h->stream));
        }
        return B200NP_OK;
    } catch (int e) { return e; }
}

int b200eb_project(b200eb_t* h, double* vel, const b200np_fab* vel_box, const double* sigma, const b200np_fab* sigma_box, double const_sigma, double* phi,
                   const b200np_fab* phi_box, double* gphi, const b200np_fab* gphi_box, double rtol, double atol, b200np_stats* stats)
{
    b200np_stats local{};
    b200np_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof(*st));
    if (!h || !vel || !h->have_geometry) return st->status = B200NP_ERR_BAD_ARG;
    if (!eb_cell_box_ok(h, vel_box, 1, 3)) return st->status = B200NP_ERR_BAD_ARG;
    if (sigma ? !eb_cell_box_ok(h, sigma_box, 0, 1) : !(const_sigma > 0)) return st->status = B200NP_ERR_BAD_ARG;
    if (phi && !eb_node_box_ok(h, phi_box)) return st->status = B200NP_ERR_BAD_ARG;
    if (gphi && !eb_cell_box_ok(h, gphi_box, 0, 3)) return st->status = B200NP_ERR_BAD_ARG;
    try {
        ECK(cudaSetDevice(h->device));
        h->launches = 0;
        ECK(cudaEventRecord(h->ev[0], h->stream));
        bool sv, ss, sp, sg;
        double* dv = eb_stage_in(h, 3, vel, vel_box, true, &sv, st);
        double* ds = eb_stage_in(h, 4, sigma, sigma_box, true, &ss, st);
        double* dp = eb_stage_in(h, 5, phi, phi_box, false, &sp, st);
        double* dg = eb_stage_in(h, 6, gphi, gphi_box, false, &sg, st);
        const int* n = h->geom.n_cell;
        ELAUNCH(h, k_eb_copy_sigma, eb_grid(h->ncell), 256, n[0], n[1], n[2], efab(ds, sigma_box), const_sigma, h->sigma);
        const int status = eb_project_dev(h, sigma ? 0.0 : const_sigma, efab(dv, vel_box), EFab{}, efab(dp, phi_box), phi_box, 0, efab(dg, gphi_box), 0, rtol, atol, st);
        eb_stage_out(h, 3, vel, vel_box, sv, st); eb_stage_out(h, 5, phi, phi_box, sp, st); eb_stage_out(h, 6, gphi, gphi_box, sg, st);
        ECK(cudaEventRecord(h->ev[1], h->stream));
        ECK(cudaEventSynchronize(h->ev[1]));
        float ms = 0;
        ECK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1])); st->ms_total = ms;
        ECK(cudaEventElapsedTime(&ms, h->ev[2], h->ev[3])); st->ms_solve = ms;
        st->launches = h->launches;
        return st->status = status;
    } catch (int e) { return st->status = e; }
}

int b200eb_apply_nodal_projection(b200eb_t* h, double* velocity, const b200np_fab* vel_box, const double* velocity_o, const double* density,
                                  const b200np_fab* rho_box, double ro_0, double* gp, const b200np_fab* gp_box, double* p_nd, const b200np_fab* p_box,
                                  const double* inflow_vel, double scaling_factor, int incremental, int proj_for_small_dt, double rtol, double atol,
                                  b200np_stats* stats)
{
    b200np_stats local{};
    b200np_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof(*st));
    if (!h || !velocity || !gp || !p_nd || !h->have_geometry) return st->status = B200NP_ERR_BAD_ARG;
    const bool use_old = incremental || proj_for_small_dt;
    if (use_old && !velocity_o) return st->status = B200NP_ERR_BAD_ARG;
    if (!eb_cell_box_ok(h, vel_box, 1, 3) || !eb_cell_box_ok(h, gp_box, 0, 3) || !eb_node_box_ok(h, p_box)) return st->status = B200NP_ERR_BAD_ARG;
    if (density ? !eb_cell_box_ok(h, rho_box, 0, 1) : !(ro_0 > 0)) return st->status = B200NP_ERR_BAD_ARG;
    try {
        ECK(cudaSetDevice(h->device));
        h->launches = 0;
        ECK(cudaEventRecord(h->ev[0], h->stream));
        bool sv, so, sr, sg, sp, si;
        double* dv = eb_stage_in(h, 3, velocity, vel_box, true, &sv, st);
        double* dvo = eb_stage_in(h, 4, velocity_o, vel_box, use_old, &so, st);
        double* dr = eb_stage_in(h, 5, density, rho_box, true, &sr, st);
        double* dgp = eb_stage_in(h, 6, gp, gp_box, true, &sg, st);
        double* dp = eb_stage_in(h, 7, p_nd, p_box, incremental != 0, &sp, st);
        const bool set_inflow = !proj_for_small_dt && !incremental;   // :81
        double* din = eb_stage_in(h, 8, set_inflow ? inflow_vel : nullptr, vel_box, true, &si, st);
        const EFab fvel = efab(dv, vel_box), fvelo = efab(use_old ? dvo : nullptr, vel_box), frho = efab(dr, rho_box), fgp = efab(dgp, gp_box),
                   fp = efab(dp, p_box), fin = efab(din, vel_box);
        const int* n = h->geom.n_cell;
        const b200np_geom& G = h->geom;
        ELAUNCH(h, k_eb_pre_add_sigma, eb_grid(h->ncell), 256, n[0], n[1], n[2], fvel, fgp, frho, fvelo, scaling_factor, ro_0, incremental ? 0 : 1, use_old ? 1 : 0,
                h->sigma);
        const long long nb = eb_box_doubles(vel_box) / std::max(vel_box->ncomp, 1);
        ELAUNCH(h, k_eb_set_vel_ghosts, eb_grid(nb), 256, n[0], n[1], n[2], vel_box->lo[0], vel_box->hi[0], vel_box->lo[1], vel_box->hi[1], vel_box->lo[2],
                vel_box->hi[2], fvel, fin, G.bc_lo[0] == B200NP_BC_INFLOW, G.bc_lo[1] == B200NP_BC_INFLOW, G.bc_lo[2] == B200NP_BC_INFLOW,
                G.bc_hi[0] == B200NP_BC_INFLOW, G.bc_hi[1] == B200NP_BC_INFLOW, G.bc_hi[2] == B200NP_BC_INFLOW);
        const int status = eb_project_dev(h, density ? 0.0 : scaling_factor / ro_0, fvel, fvelo, fp, p_box, incremental ? 1 : 0, fgp, incremental ? 1 : 0, rtol, atol, st);
        eb_stage_out(h, 3, velocity, vel_box, sv, st); eb_stage_out(h, 6, gp, gp_box, sg, st); eb_stage_out(h, 7, p_nd, p_box, sp, st);
        ECK(cudaEventRecord(h->ev[1], h->stream));
        ECK(cudaEventSynchronize(h->ev[1]));
        float ms = 0;
        ECK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1])); st->ms_total = ms;
        ECK(cudaEventElapsedTime(&ms, h->ev[2], h->ev[3])); st->ms_solve = ms;
        st->launches = h->launches;
        return st->status = status;
    } catch (int e) { return st->status = e; }
}

}  // extern "C"

namespace {
struct EbMfView {
    const b200np_mfab* m = nullptr;
    int slot = 0, ncomp = 1;
    bool staged = false;
    std::vector<size_t> off;
    double* dense = nullptr;
    b200np_fab box{};        // the dense box
};
size_t ebmf_fab_doubles(const b200np_fab& b, int ncomp)
{
    return (size_t)(b.hi[0] - b.lo[0] + 1) * (b.hi[1] - b.lo[1] + 1) * (b.hi[2] - b.lo[2] + 1) * ncomp;
}
// the valid boxes (allocated box shrunk by ngrow) must lie inside [lo, hi] and, for cell-centred MultiFabs, tile it
bool ebmf_ok(const b200np_mfab* m, const int lo[3], const int hi[3], int ncomp, bool tile, int min_grow = 0)
{
    if (!m || m->nfabs < 1 || m->ngrow < min_grow || m->ncomp < ncomp || !m->box || !m->data) return false;
    long long vol = 0;
    for (int f = 0; f < m->nfabs; ++f) {
        if (!m->data[f]) return false;
        long long v = 1;
        for (int d = 0; d < 3; ++d) {
            const int vlo = m->box[f].lo[d] + m->ngrow, vhi = m->box[f].hi[d] - m->ngrow;
            if (vhi < vlo || vlo < lo[d] || vhi > hi[d]) return false;
            v *= vhi - vlo + 1;
        }
        vol += v;
    }
    if (tile) {
        long long want = 1;
        for (int d = 0; d < 3; ++d) want *= hi[d] - lo[d] + 1;
        if (vol != want) return false;
    }
    return true;
}
// fabs on the device (staged when they are host pointers) + a dense array over `box`
void ebmf_map(b200eb* h, EbMfView& V, int slot, const b200np_mfab* m, int ncomp, const b200np_fab& box, bool copy_in, b200np_stats* st)
{
    V = EbMfView{};
    if (!m) return;
    V.m = m; V.slot = slot; V.ncomp = ncomp; V.box = box; V.box.ncomp = ncomp;
    auto& S = h->mf[slot];
    const int nf = m->nfabs;
    V.staged = !eb_is_dev_ptr(m->data[0]);
    V.off.assign(nf, 0);
    size_t total = 0;
    std::vector<EbMfFab> host(nf);
    for (int f = 0; f < nf; ++f) {
        const b200np_fab& b = m->box[f];
        host[f].nx = b.hi[0] - b.lo[0] + 1; host[f].ny = b.hi[1] - b.lo[1] + 1; host[f].nz = b.hi[2] - b.lo[2] + 1;
        host[f].cs = (long long)host[f].nx * host[f].ny * host[f].nz;
        for (int q = 0; q < 3; ++q) host[f].lo[q] = b.lo[q];
        V.off[f] = total;
        total += ebmf_fab_doubles(b, m->ncomp);
    }
    if (V.staged) {
        if (S.stage_doubles < total) { if (S.stage) ECK(cudaFree(S.stage)); ECK(cudaMalloc(&S.stage, total * sizeof(double))); S.stage_doubles = total; }
        if (copy_in) {
            for (int f = 0; f < nf; ++f)
                ECK(cudaMemcpyAsync(S.stage + V.off[f], m->data[f], ebmf_fab_doubles(m->box[f], m->ncomp) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            if (st) st->h2d_bytes += (long long)(total * sizeof(double));
        }
    }
    for (int f = 0; f < nf; ++f) host[f].p = V.staged ? S.stage + V.off[f] : m->data[f];
    if (S.tab_cap < (size_t)nf) { if (S.tab) ECK(cudaFree(S.tab)); ECK(cudaMalloc(&S.tab, (size_t)nf * sizeof(EbMfFab))); S.tab_cap = nf; }
    ECK(cudaMemcpyAsync(S.tab, host.data(), (size_t)nf * sizeof(EbMfFab), cudaMemcpyHostToDevice, h->stream));
    ECK(cudaStreamSynchronize(h->stream));   // `host` goes out of scope
    const size_t nd = ebmf_fab_doubles(box, ncomp);
    if (S.dense_doubles < nd) { if (S.dense) ECK(cudaFree(S.dense)); ECK(cudaMalloc(&S.dense, nd * sizeof(double))); S.dense_doubles = nd; }
    V.dense = S.dense;
}
EDense ebmf_dense(const EbMfView& V)
{
    EDense d{};
    d.p = V.dense;
    for (int q = 0; q < 3; ++q) { d.lo[q] = V.box.lo[q]; d.hi[q] = V.box.hi[q]; }
    d.cs = (long long)(V.box.hi[0] - V.box.lo[0] + 1) * (V.box.hi[1] - V.box.lo[1] + 1) * (V.box.hi[2] - V.box.lo[2] + 1);
    return d;
}
dim3 ebmf_grid(const b200np_mfab* m)
{
    long long mx = 1;
    for (int f = 0; f < m->nfabs; ++f) mx = std::max<long long>(mx, (long long)ebmf_fab_doubles(m->box[f], 1));
    return dim3((unsigned)std::min<long long>((mx + 255) / 256, 64), (unsigned)m->nfabs);
}
void ebmf_gather(b200eb* h, EbMfView& V, int mode)
{
    if (!V.m) return;
    const int* n = h->geom.n_cell;
    ELAUNCH(h, k_eb_mf_gather, ebmf_grid(V.m), 256, (const EbMfFab*)h->mf[V.slot].tab, V.m->ngrow, V.ncomp, ebmf_dense(V), mode, n[0], n[1], n[2]);
}
void ebmf_scatter(b200eb* h, EbMfView& V, int mode, b200np_stats* st)
{
    if (!V.m) return;
    ELAUNCH(h, k_eb_mf_scatter, ebmf_grid(V.m), 256, (const EbMfFab*)h->mf[V.slot].tab, V.m->ngrow, V.ncomp, ebmf_dense(V), mode);
    if (!V.staged) return;
    size_t total = 0;
    for (int f = 0; f < V.m->nfabs; ++f) {
        const size_t nd = ebmf_fab_doubles(V.m->box[f], V.m->ncomp);
        ECK(cudaMemcpyAsync(V.m->data[f], h->mf[V.slot].stage + V.off[f], nd * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        total += nd;
    }
    if (st) st->d2h_bytes += (long long)(total * sizeof(double));
}
void ebmf_boxes(const b200eb* h, b200np_fab& cells, b200np_fab& grown, b200np_fab& nodes)
{
    for (int d = 0; d < 3; ++d) {
        cells.lo[d] = 0; cells.hi[d] = h->geom.n_cell[d] - 1;
        grown.lo[d] = -1; grown.hi[d] = h->geom.n_cell[d];
        nodes.lo[d] = 0; nodes.hi[d] = h->geom.n_cell[d];
    }
    cells.ncomp = 1; grown.ncomp = 3; nodes.ncomp = 1;
}
}  // namespace

extern "C" {

// b200eb_set_geometry over multi-box MultiFabs (getVolFrac(), the integral MultiFab)
int b200eb_set_geometry_mf(b200eb_t* h, const b200np_mfab* vfrac, const b200np_mfab* intg)
{
    if (!h) return B200NP_ERR_BAD_ARG;
    b200np_fab cells{}, grown{}, nodes{};
    ebmf_boxes(h, cells, grown, nodes);
    if (!ebmf_ok(vfrac, cells.lo, cells.hi, 1, true) || !ebmf_ok(intg, cells.lo, cells.hi, 18, true)) return B200NP_ERR_BAD_ARG;
    try {
        ECK(cudaSetDevice(h->device));
        EbMfView Vv, Vi;
        ebmf_map(h, Vv, 0, vfrac, 1, cells, true, nullptr);
        ebmf_map(h, Vi, 1, intg, 18, cells, true, nullptr);
        ebmf_gather(h, Vv, EBMF_VALID);
        ebmf_gather(h, Vi, EBMF_VALID);
        ECK(cudaStreamSynchronize(h->stream));
        return b200eb_set_geometry(h, Vv.dense, &Vv.box, Vi.dense, &Vi.box);
    } catch (int e) { return e; }
}

// b200eb_project over multi-box MultiFabs: bit-identical to the single-box call.  After the call every cell of vel's fabs that lies inside the
// domain grown by one cell holds the projected velocity / the BC ghost value (interior ghost cells as after FillBoundary).
int b200eb_project_mf(b200eb_t* h, const b200np_mfab* vel, const b200np_mfab* sigma, double const_sigma, const b200np_mfab* phi, const b200np_mfab* gphi,
                      double rtol, double atol, b200np_stats* stats)
{
    b200np_stats local{};
    b200np_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof(*st));
    if (!h || !h->have_geometry) return st->status = B200NP_ERR_BAD_ARG;
    b200np_fab cells{}, grown{}, nodes{};
    ebmf_boxes(h, cells, grown, nodes);
    if (!ebmf_ok(vel, cells.lo, cells.hi, 3, true, 1)) return st->status = B200NP_ERR_BAD_ARG;
    if (sigma && !ebmf_ok(sigma, cells.lo, cells.hi, 1, true)) return st->status = B200NP_ERR_BAD_ARG;
    if (gphi && !ebmf_ok(gphi, cells.lo, cells.hi, 3, true)) return st->status = B200NP_ERR_BAD_ARG;
    if (phi && !ebmf_ok(phi, nodes.lo, nodes.hi, 1, false)) return st->status = B200NP_ERR_BAD_ARG;
    try {
        ECK(cudaSetDevice(h->device));
        b200np_stats outer{};
        EbMfView Vv, Vs, Vp, Vg;
        ebmf_map(h, Vv, 2, vel, 3, grown, true, &outer);
        ebmf_map(h, Vs, 3, sigma, 1, cells, true, &outer);
        ebmf_map(h, Vp, 4, phi, 1, nodes, false, &outer);
        ebmf_map(h, Vg, 5, gphi, 3, cells, false, &outer);
        ECK(cudaMemsetAsync(Vv.dense, 0, ebmf_fab_doubles(grown, 3) * sizeof(double), h->stream));
        ebmf_gather(h, Vv, EBMF_VALID_BC);
        ebmf_gather(h, Vs, EBMF_VALID);
        ECK(cudaStreamSynchronize(h->stream));
        const int rc = b200eb_project(h, Vv.dense, &Vv.box, sigma ? Vs.dense : nullptr, sigma ? &Vs.box : nullptr, const_sigma, phi ? Vp.dense : nullptr,
                                      phi ? &Vp.box : nullptr, gphi ? Vg.dense : nullptr, gphi ? &Vg.box : nullptr, rtol, atol, st);
        if (rc != B200NP_OK && rc != B200NP_ERR_NOT_CONVERGED && rc != B200NP_ERR_DIVERGED) return rc;
        ebmf_scatter(h, Vv, EBMF_VALID_BC, &outer);
        ebmf_scatter(h, Vp, EBMF_VALID, &outer);
        ebmf_scatter(h, Vg, EBMF_VALID, &outer);
        ECK(cudaStreamSynchronize(h->stream));
        st->h2d_bytes += outer.h2d_bytes; st->d2h_bytes += outer.d2h_bytes;
        return st->status = rc;
    } catch (int e) { return st->status = e; }
}

// incflo::ApplyNodalProjection under AMREX_USE_EB over multi-box LevelData MultiFabs
int b200eb_apply_nodal_projection_mf(b200eb_t* h, const b200np_mfab* velocity, const b200np_mfab* velocity_o, const b200np_mfab* density, double ro_0,
                                     const b200np_mfab* gp, const b200np_mfab* p_nd, const b200np_mfab* inflow_vel, double scaling_factor, int incremental,
                                     int proj_for_small_dt, double rtol, double atol, b200np_stats* stats)
{
    b200np_stats local{};
    b200np_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof(*st));
    if (!h || !h->have_geometry) return st->status = B200NP_ERR_BAD_ARG;
    const bool use_old = incremental || proj_for_small_dt;
    const bool set_inflow = !proj_for_small_dt && !incremental;
    b200np_fab cells{}, grown{}, nodes{};
    ebmf_boxes(h, cells, grown, nodes);
    if (!ebmf_ok(velocity, cells.lo, cells.hi, 3, true, 1)) return st->status = B200NP_ERR_BAD_ARG;
    if (use_old && !ebmf_ok(velocity_o, cells.lo, cells.hi, 3, true)) return st->status = B200NP_ERR_BAD_ARG;
    if (density && !ebmf_ok(density, cells.lo, cells.hi, 1, true)) return st->status = B200NP_ERR_BAD_ARG;
    if (!ebmf_ok(gp, cells.lo, cells.hi, 3, true) || !ebmf_ok(p_nd, nodes.lo, nodes.hi, 1, false)) return st->status = B200NP_ERR_BAD_ARG;
    if (inflow_vel && set_inflow && !ebmf_ok(inflow_vel, cells.lo, cells.hi, 3, true, 1)) return st->status = B200NP_ERR_BAD_ARG;
    try {
        ECK(cudaSetDevice(h->device));
        b200np_stats outer{};
        EbMfView Vv, Vo, Vr, Vg, Vp, Vi;
        b200np_fab c3 = cells; c3.ncomp = 3;
        ebmf_map(h, Vv, 2, velocity, 3, grown, true, &outer);
        ebmf_map(h, Vo, 3, use_old ? velocity_o : nullptr, 3, grown, true, &outer);
        ebmf_map(h, Vr, 4, density, 1, cells, true, &outer);
        ebmf_map(h, Vg, 5, gp, 3, c3, true, &outer);
        ebmf_map(h, Vp, 6, p_nd, 1, nodes, incremental != 0, &outer);
        ebmf_map(h, Vi, 7, (inflow_vel && set_inflow) ? inflow_vel : nullptr, 3, grown, true, &outer);
        ECK(cudaMemsetAsync(Vv.dense, 0, ebmf_fab_doubles(grown, 3) * sizeof(double), h->stream));
        ebmf_gather(h, Vv, EBMF_VALID);       // ghost cells are zeroed by the call itself (setBndry(0)); only the inflow layer comes from inflow_vel
        if (Vo.m) { ECK(cudaMemsetAsync(Vo.dense, 0, ebmf_fab_doubles(grown, 3) * sizeof(double), h->stream)); ebmf_gather(h, Vo, EBMF_VALID); }
        ebmf_gather(h, Vr, EBMF_VALID);
        ebmf_gather(h, Vg, EBMF_VALID);
        if (incremental) ebmf_gather(h, Vp, EBMF_VALID);
        if (Vi.m) { ECK(cudaMemsetAsync(Vi.dense, 0, ebmf_fab_doubles(grown, 3) * sizeof(double), h->stream)); ebmf_gather(h, Vi, EBMF_VALID_BC); }
        ECK(cudaStreamSynchronize(h->stream));
        const int rc = b200eb_apply_nodal_projection(h, Vv.dense, &Vv.box, Vo.m ? Vo.dense : nullptr, density ? Vr.dense : nullptr, density ? &Vr.box : nullptr, ro_0,
                                                     Vg.dense, &Vg.box, Vp.dense, &Vp.box, Vi.m ? Vi.dense : nullptr, scaling_factor, incremental, proj_for_small_dt,
                                                     rtol, atol, st);
        if (rc != B200NP_OK && rc != B200NP_ERR_NOT_CONVERGED && rc != B200NP_ERR_DIVERGED) return rc;
        ebmf_scatter(h, Vv, EBMF_VALID_BC, &outer);
        ebmf_scatter(h, Vg, EBMF_VALID, &outer);
        ebmf_scatter(h, Vp, EBMF_VALID, &outer);
        ECK(cudaStreamSynchronize(h->stream));
        st->h2d_bytes += outer.h2d_bytes; st->d2h_bytes += outer.d2h_bytes;
        return st->status = rc;
    } catch (int e) { return st->status = e; }
}

int b200eb_level_dims(const b200eb_t* h, int lev, int n_cell[3], int n_node[3])
{
    if (!h || lev < 0 || lev >= (int)h->lv.size()) return B200NP_ERR_BAD_ARG;
    for (int d = 0; d < 3; ++d) { n_cell[d] = h->lv[lev].g.n[d]; n_node[d] = h->lv[lev].g.nn[d]; }
    return B200NP_OK;
}

// test hook: build the stencil hierarchy from sigma (host or device cell array, or NULL + const_sigma) without projecting
int b200eb_build_stencils(b200eb_t* h, const double* sigma, const b200np_fab* sigma_box, double const_sigma)
{
    if (!h || !h->have_geometry) return B200NP_ERR_BAD_ARG;
    if (sigma ? !eb_cell_box_ok(h, sigma_box, 0, 1) : !(const_sigma > 0)) return B200NP_ERR_BAD_ARG;
    try {
        ECK(cudaSetDevice(h->device));
        bool ss;
        double* ds = eb_stage_in(h, 4, sigma, sigma_box, true, &ss, nullptr);
        const int* n = h->geom.n_cell;
        ELAUNCH(h, k_eb_copy_sigma, eb_grid(h->ncell), 256, n[0], n[1], n[2], efab(ds, sigma_box), const_sigma, h->sigma);
        eb_build_stencils(h, sigma ? 0.0 : const_sigma);
        ECK(cudaStreamSynchronize(h->stream));
        return B200NP_OK;
    } catch (int e) { return e; }
}

// test hook: the 13 forward entries + diagonal of level lev, host array (14, nnz, nny, nnx)
int b200eb_level_stencil(b200eb_t* h, int lev, double* out)
{
    if (!h || lev < 0 || lev >= (int)h->lv.size() || !out || !h->have_stencil) return B200NP_ERR_BAD_ARG;
    try {
        ECK(cudaSetDevice(h->device));
        EbLevel& L = h->lv[lev];
        ELAUNCH(h, k_eb_export_stencil, eb_grid(L.g.nnode), 256, L.g, h->tmp_nat);
        ECK(cudaMemcpyAsync(out, h->tmp_nat, 14 * eb_nreal(L.g) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        ECK(cudaStreamSynchronize(h->stream));
        return B200NP_OK;
    } catch (int e) { return e; }
}

// test hooks on level arrays; host arrays in natural node order (nnz, nny, nnx):
// op 0 smooth (arg MLMG smooth calls: x = in_a, rhs = in_b), 1 residual (in_b - A in_a), 2 restriction of in_a to level lev + 1,
// 3 in_a + interpolation of in_b (level lev + 1), 4 bottom solve of in_b on the coarsest level, 5 A in_a
int b200eb_level_op(b200eb_t* h, int lev, int op, int arg, const double* in_a, const double* in_b, double* out)
{
    if (!h || lev < 0 || lev >= (int)h->lv.size() || !h->have_stencil || !out) return B200NP_ERR_BAD_ARG;
    try {
        ECK(cudaSetDevice(h->device));
        EbLevel& L = h->lv[lev];
        auto up = [&](EbLevel& T, double* d, const double* src) {
            if (!src) return;
            ECK(cudaMemcpyAsync(h->tmp_nat, src, eb_nreal(T.g) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            ELAUNCH(h, k_eb_permute, eb_grid(T.g.nnode), 256, T.g, (const double*)h->tmp_nat, d, 0);
        };
        auto down = [&](EbLevel& T, const double* d) {
            ELAUNCH(h, k_eb_permute, eb_grid(T.g.nnode), 256, T.g, d, h->tmp_nat, 1);
            ECK(cudaMemcpyAsync(out, h->tmp_nat, eb_nreal(T.g) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        };
        switch (op) {
        case 0:
            up(L, L.cor, in_a); up(L, L.res, in_b);
            eb_smooth(h, L, L.cor, L.res, arg);
            down(L, L.cor);
            break;
        case 1:
            up(L, L.cor, in_a); up(L, L.res, in_b);
            eb_launch_residual(h, L, (const double*)L.cor, (const double*)L.res, L.rescor, (double*)nullptr);
            down(L, L.rescor);
            break;
        case 2: {
            if (lev + 1 >= (int)h->lv.size()) return B200NP_ERR_BAD_ARG;
            EbLevel& C = h->lv[lev + 1];
            up(L, L.rescor, in_a);
            ELAUNCH(h, k_eb_restrict, eb_grid(C.g.nnode), 256, C.g, L.g, (const double*)L.rescor, C.res);
            down(C, C.res);
            break;
        }
        case 3: {
            if (lev + 1 >= (int)h->lv.size()) return B200NP_ERR_BAD_ARG;
            EbLevel& C = h->lv[lev + 1];
            up(L, L.cor, in_a); up(C, C.cor, in_b);
            ELAUNCH(h, k_eb_interp_add, eb_grid(L.g.nnode), 256, L.g, C.g, L.cor, (const double*)C.cor);
            down(L, L.cor);
            break;
        }
        case 4: {
            EbLevel& B = h->lv.back();
            up(B, B.res, in_b);
            ECK(cudaMemsetAsync(h->dinfo, 0, 4 * sizeof(int), h->stream));
            eb_bottom(h);
            down(B, B.cor);
            break;
        }
        case 5: {   // A x = -(0 - A x)
            up(L, L.cor, in_a);
            ECK(cudaMemsetAsync(L.res, 0, L.g.nnode * sizeof(double), h->stream));
            eb_launch_residual(h, L, (const double*)L.cor, (const double*)L.res, L.rescor, (double*)nullptr);
            ELAUNCH(h, k_eb_permute, eb_grid(L.g.nnode), 256, L.g, (const double*)L.rescor, h->tmp_nat, 1);
            ECK(cudaMemcpyAsync(out, h->tmp_nat, eb_nreal(L.g) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            ECK(cudaStreamSynchronize(h->stream));
            for (size_t t = 0; t < eb_nreal(L.g); ++t) out[t] = -out[t];
            return B200NP_OK;
        }
        default: return B200NP_ERR_BAD_ARG;
        }
        ECK(cudaStreamSynchronize(h->stream));
        return B200NP_OK;
    } catch (int e) { return e; }
}

// measurement hook: `reps` times { arg smooth calls (op 0) | one residual (op 1) } on the device arrays of level lev, timed with CUDA
// events on the handle's stream; ms = average per repetition.  Needs built stencils.
int b200eb_time_op(b200eb_t* h, int lev, int op, int arg, int reps, double* ms)
{
    if (!h || lev < 0 || lev >= (int)h->lv.size() || !h->have_stencil || !ms || reps < 1) return B200NP_ERR_BAD_ARG;
    try {
        ECK(cudaSetDevice(h->device));
        EbLevel& L = h->lv[lev];
        for (int pass = 0; pass < 2; ++pass) {   // pass 0: warm-up
            ECK(cudaEventRecord(h->ev[0], h->stream));
            for (int r = 0; r < (pass ? reps : 1); ++r) {
                if (op == 0) eb_smooth(h, L, L.cor, L.res, arg);
                else if (op == 1) {
                    eb_launch_residual(h, L, (const double*)L.cor, (const double*)L.res, L.rescor, (double*)nullptr);
                } else return B200NP_ERR_BAD_ARG;
            }
            ECK(cudaEventRecord(h->ev[1], h->stream));
            ECK(cudaEventSynchronize(h->ev[1]));
        }
        float t = 0;
        ECK(cudaEventElapsedTime(&t, h->ev[0], h->ev[1]));
        *ms = (double)t / reps;
        return B200NP_OK;
    } catch (int e) { return e; }
}

// test hook: rhs = D u (+ EB inflow) of the caller's velocity (one ghost layer), natural node order; needs built stencils (active set)
int b200eb_compute_rhs(b200eb_t* h, const double* vel, const b200np_fab* vel_box, double* out)
{
    if (!h || !vel || !out || !h->have_stencil || !eb_cell_box_ok(h, vel_box, 1, 3)) return B200NP_ERR_BAD_ARG;
    try {
        ECK(cudaSetDevice(h->device));
        bool sv;
        double* dv = eb_stage_in(h, 3, vel, vel_box, true, &sv, nullptr);
        EbLevel& L0 = h->lv[0];
        const b200np_geom& G = h->geom;
        ELAUNCH(h, k_eb_divu, eb_grid(L0.g.nnode, 128), 128, L0.g, (const double*)h->geo, efab(dv, vel_box), (const double*)(h->have_ebflow ? h->ebf : nullptr),
                1.0 / G.dx[0], 1.0 / G.dx[1], 1.0 / G.dx[2], L0.rhs);
        ELAUNCH(h, k_eb_permute, eb_grid(L0.g.nnode), 256, L0.g, (const double*)L0.rhs, h->tmp_nat, 1);
        ECK(cudaMemcpyAsync(out, h->tmp_nat, eb_nreal(L0.g) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        ECK(cudaStreamSynchronize(h->stream));
        return B200NP_OK;
    } catch (int e) { return e; }
}

}  // extern "C"
