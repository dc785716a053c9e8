// np_level.h -- device-visible description of one multigrid level and the index maps
// shared by every kernel.
//
// Data layout in HBM:
//   * nodal fields are stored once per UNIQUE node: nn[d] = n[d] in a periodic direction,
//     n[d]+1 otherwise; i fastest with pitch px (multiple of 8 doubles = 64 B), then j, then
//     the locally owned z planes [k0, k0+nzl) plus one ghost plane slot on either side
//     (slot -1 and slot nzl).  Ghost slots are only used by the slab-decomposed multi-GPU
//     path; on one GPU the z neighbours are found by wrap / reflection like x and y.
//   * cell fields (sigma): n[d] cells, pitch cpx, same ghost-plane-slot convention.
// Boundary conditions follow MLNodeLinOp::applyBC: periodic wrap, Neumann /
// inflow reflection phi(-1) = phi(1), Dirichlet nodes masked to 0; sigma ghost cells copy the
// adjacent interior cell.
#pragma once
#include <cuda_runtime.h>

namespace b200np_dev {

struct Lev {
    int n[3];    // global cells
    int nn[3];   // global unique nodes
    int per[3];  // periodic
    int rlo[3], rhi[3];  // Neumann / inflow (reflecting) face
    int dlo[3], dhi[3];  // Dirichlet face
    int px;              // node pitch in x (doubles)
    long long ps;        // node plane stride
    int cpx;             // cell pitch in x
    long long cps;       // cell plane stride
    int k0, nzl;         // owned node planes: global [k0, k0+nzl)
    int ck0, cnzl;       // owned cell planes
    int dist;            // 1: z neighbours live in ghost plane slots (multi-GPU slab)
    double dxinv[3];
    // stencil factors of mlndlap_adotx_aa
    double fxyz, fmx2y2z, f2xmy2z, f2x2ymz, f4xm2ym2z, fm2x4ym2z, fm2xm2y4z;
    double csig;          // constant sigma (used when sigma == nullptr)
    const double* sigma;  // cell array, plane 0 of the owned range (ghost slot at -1)
    // incflo BC::mixed faces: the outflow part of the face is Dirichlet through the overset mask of
    // incflo::make_nodalBC_mask (src/boundary_conditions/incflo_set_bcs.cpp:10-53, prob_set_BC_MF src/prob/prob_bc.cpp:9-101):
    // bit (dir + 3 side) of mixm marks a mixed face; on a low-side face the nodes with idx[mixdir] <= mixhalf are
    // masked, on a high-side face those with idx[mixdir] > mixhalf.  mixhalf = (domain.length(mixdir) / 2) >> level:
    // a coarser multigrid level takes the mask of its nodes from the finer level by injection (node 2i).
    int mixm, mixdir, mixhalf;
};

// Programmatic dependent launch (PDL).  pdl_wait(): block until the predecessor kernel has completed
// and its writes are visible -- must precede the first access to data it produced.
// pdl_trigger(): allow the successor to become resident (it still waits in its own pdl_wait()).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// small grids (a single wave) let the successor in right away; multi-wave grids only at their tail,
// so that waiting successor CTAs do not take SM slots from this grid's later waves
__device__ __forceinline__ bool pdl_small_grid() { return gridDim.x * gridDim.y * gridDim.z <= 296u; }

// ---- NVLink peer-memory halo flags (slab-decomposed path; protocol in np_kernels.cuh K10) ----
struct HaloFlags {
    unsigned long long* my;       // [0] written by my lower neighbour, [1] by my upper neighbour, [2] epoch base,
                                  // [4], [5] counters of the fused sweeps (np_smooth3.cuh), [6] a wait timed out (halo_spin)
    unsigned long long* lo_flag;  // lower neighbour's word [1] (peer pointer) or nullptr
    unsigned long long* hi_flag;  // upper neighbour's word [0] (peer pointer) or nullptr
    unsigned long long k;         // this exchange is number k since the base was last advanced: epoch = my[2] + 1 + k
};
__device__ __forceinline__ unsigned long long ld_relaxed_gpu(const unsigned long long* p);
__device__ __forceinline__ unsigned long long halo_epoch(const HaloFlags& f) { return ld_relaxed_gpu(f.my + 2) + 1ull + f.k; }
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_gpu(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Bounded wait for a neighbour's flag word my[which] to reach epoch e.  A rank that died, or left the call early with an
// error of its own, must not hang this GPU for ever: after 2^33 cycles (about 4 s) the wait gives up and raises my[6]; the
// host turns that into B200NP_ERR_PEER_TIMEOUT at the end of the solve (the numbers of such a solve are garbage).
__device__ __forceinline__ void halo_spin(unsigned long long* my, int which, unsigned long long e)
{
    const long long t0 = clock64();
    while (ld_acquire_sys(my + which) < e) {
        __nanosleep(20);
        if (clock64() - t0 > (1ll << 33)) { atomicExch(my + 6, 1ull); break; }
    }
}

// node index in [-1, n+1] -> unique storage index
__host__ __device__ __forceinline__ int nmap(int i, int n, int per)
{
    if (per) { if (i < 0) i += n; else if (i >= n) i -= n; }
    else     { if (i < 0) i = -i; else if (i > n) i = 2 * n - i; }
    return i;
}
// cell index in [-1, n] -> storage index
__host__ __device__ __forceinline__ int cmap(int i, int n, int per)
{
    if (per) { if (i < 0) i += n; else if (i >= n) i -= n; }
    else     { if (i < 0) i = 0; else if (i >= n) i = n - 1; }
    return i;
}
// local node plane index (may be -1 or nzl) -> local storage plane
__host__ __device__ __forceinline__ int zplane(const Lev& L, int kl)
{
    return L.dist ? kl : nmap(kl, L.n[2], L.per[2]);
}
__host__ __device__ __forceinline__ int czplane(const Lev& L, int kl)
{
    return L.dist ? kl : cmap(kl, L.n[2], L.per[2]);
}
// does the level have any Dirichlet (masked) node at all?
__host__ __device__ __forceinline__ bool lev_any_masked(const Lev& L)
{
    return (L.dlo[0] | L.dhi[0] | L.dlo[1] | L.dhi[1] | L.dlo[2] | L.dhi[2] | L.mixm) != 0;
}
__host__ __device__ __forceinline__ bool node_masked(const Lev& L, int i, int j, int kg)
{
    bool m = (L.dlo[0] && i == 0) || (L.dhi[0] && i == L.n[0]) || (L.dlo[1] && j == 0) || (L.dhi[1] && j == L.n[1]) ||
             (L.dlo[2] && kg == 0) || (L.dhi[2] && kg == L.n[2]);
    if (L.mixm) {
        const int t = L.mixdir == 0 ? i : L.mixdir == 1 ? j : kg;
        const bool lo = t <= L.mixhalf;
        m = m || ((L.mixm & 1) && i == 0 && lo) || ((L.mixm & 8) && i == L.n[0] && !lo) ||
            ((L.mixm & 2) && j == 0 && lo) || ((L.mixm & 16) && j == L.n[1] && !lo) ||
            ((L.mixm & 4) && kg == 0 && lo) || ((L.mixm & 32) && kg == L.n[2] && !lo);
    }
    return m;
}
// dot-product / solvability weight: 1/2 per reflecting boundary direction
__host__ __device__ __forceinline__ double node_weight(const Lev& L, int i, int j, int kg)
{
    if (node_masked(L, i, j, kg)) return 0.0;
    double w = 1.0;
    if (!L.per[0]) { if (L.rlo[0] && i == 0) w *= 0.5; if (L.rhi[0] && i == L.n[0]) w *= 0.5; }
    if (!L.per[1]) { if (L.rlo[1] && j == 0) w *= 0.5; if (L.rhi[1] && j == L.n[1]) w *= 0.5; }
    if (!L.per[2]) { if (L.rlo[2] && kg == 0) w *= 0.5; if (L.rhi[2] && kg == L.n[2]) w *= 0.5; }
    return w;
}

// y = L phi at one node from the 8 surrounding sigma and the 27 phi values; returns the
// diagonal in s0.  S[c][b][a] = sigma(i-1+a, j-1+b, k-1+c); P[c][b][a] = phi(i-1+a, j-1+b, k-1+c).
// Restates mlndlap_adotx_aa; verified == Q1 finite-element stiffness by the
// oracle tests.
__device__ __forceinline__ double stencil27(const Lev& L, const double (&S)[2][2][2], const double (&P)[3][3][3],
                                            double& s0)
{
    double sumS = 0, corner = 0, ex = 0, ey = 0, ez = 0, fxs = 0, fys = 0, fzs = 0;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                sumS += S[c][b][a];
                corner += S[c][b][a] * P[2 * c][2 * b][2 * a];
            }
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            ex += (S[c][b][0] + S[c][b][1]) * P[2 * c][2 * b][1];   // neighbours (0, +-1, +-1)
            ey += (S[c][0][b] + S[c][1][b]) * P[2 * c][1][2 * b];   // (+-1, 0, +-1)
            ez += (S[0][c][b] + S[1][c][b]) * P[1][2 * c][2 * b];   // (+-1, +-1, 0)
        }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        fxs += (S[0][0][a] + S[0][1][a] + S[1][0][a] + S[1][1][a]) * P[1][1][2 * a];  // (+-1,0,0)
        fys += (S[0][a][0] + S[0][a][1] + S[1][a][0] + S[1][a][1]) * P[1][2 * a][1];  // (0,+-1,0)
        fzs += (S[a][0][0] + S[a][0][1] + S[a][1][0] + S[a][1][1]) * P[2 * a][1][1];  // (0,0,+-1)
    }
    s0 = -4.0 * L.fxyz * sumS;
    return s0 * P[1][1][1] + L.fxyz * corner + L.fmx2y2z * ex + L.f2xmy2z * ey + L.f2x2ymz * ez + L.f4xm2ym2z * fxs +
           L.fm2x4ym2z * fys + L.fm2xm2y4z * fzs;
}

// constant-sigma variant (mlndlap_adotx_c): S == sig everywhere
__device__ __forceinline__ double stencil27_c(const Lev& L, double sig, const double (&P)[3][3][3], double& s0)
{
    double corner = 0, ex = 0, ey = 0, ez = 0;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            corner += P[2 * c][2 * b][0] + P[2 * c][2 * b][2];
            ex += P[2 * c][2 * b][1];
            ey += P[2 * c][1][2 * b];
            ez += P[1][2 * c][2 * b];
        }
    double fxs = P[1][1][0] + P[1][1][2], fys = P[1][0][1] + P[1][2][1], fzs = P[0][1][1] + P[2][1][1];
    s0 = -32.0 * L.fxyz * sig;
    return s0 * P[1][1][1] +
           sig * (L.fxyz * corner + 2.0 * (L.fmx2y2z * ex + L.f2xmy2z * ey + L.f2x2ymz * ez) +
                  4.0 * (L.f4xm2ym2z * fxs + L.fm2x4ym2z * fys + L.fm2xm2y4z * fzs));
}

}  // namespace b200np_dev
