// np_kernels.cuh -- sm_100a kernels of the nodal projection (one GPU's slab of one level each).
//
// Kernel inventory (reference function replaced -> kernel), SURVEY.md 2.4 / DESIGN.md section 4:
//   K1  incflo_apply_nodal_projection.cpp:53-59,:68,:115-118      -> k_pre_add_sigma
//   K2  mlndlap_divu + mlndlap_impose_neumann_bc (A.2)            -> k_divu
//   K3  mlndlap_adotx_{aa,c} / solutionResidual (A.3)             -> k_residual_iso / k_residual_v2 (np_smooth*.cuh)
//   K4  mlndlap_gauss_seidel_{aa,c} / mlndlap_gscolor (A.4)       -> k_smooth_iso* / k_smooth_v2 (np_smooth*.cuh)
//   K5  mlndlap_restriction (A.5)                                 -> k_restrict
//   K6  mlndlap_interpadd_{aa,c} (A.6)                            -> k_interp_tile (np_smooth.cuh), Interp (node form)
//   K11 the levels below ~17^3 nodes, all of the above in one CTA -> k_coarse_tail (np_tail.cuh)
//   K7  mlndlap_mknewu{,_c} + copy-out :221-256 (A.7)             -> k_mknewu, k_copy_phi
//   K8  MLCGSolver BiCGStab / CG (A.10)                           -> k_bottom_bicgstab (one CTA)
//   K9  average_down of sigma (A.8)                               -> k_coarsen_sigma
#pragma once
#include "np_level.h"

namespace b200np_dev {

#define NP_TX 64
#define NP_TY 16

// ------------------------------------------------------------------------------------------
// generic gathers from global memory (used by the small / non-critical kernels)
// ------------------------------------------------------------------------------------------
template <bool VAR>
__device__ __forceinline__ void gather_sigma_g(const Lev& L, int i, int j, int kl, double (&S)[2][2][2])
{
    if (!VAR) return;
    int ci[2] = {cmap(i - 1, L.n[0], L.per[0]), cmap(i, L.n[0], L.per[0])};
    int cj[2] = {cmap(j - 1, L.n[1], L.per[1]), cmap(j, L.n[1], L.per[1])};
    int ck[2] = {czplane(L, kl - 1), czplane(L, kl)};
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int a = 0; a < 2; ++a) S[c][b][a] = L.sigma[ck[c] * L.cps + (long long)cj[b] * L.cpx + ci[a]];
}

__device__ __forceinline__ void gather_phi_g(const Lev& L, const double* __restrict__ phi, int i, int j, int kl,
                                             double (&P)[3][3][3])
{
    int ix[3], jy[3], kz[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        ix[a] = nmap(i - 1 + a, L.n[0], L.per[0]);
        jy[a] = nmap(j - 1 + a, L.n[1], L.per[1]);
        kz[a] = zplane(L, kl - 1 + a);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
            for (int a = 0; a < 3; ++a) P[c][b][a] = phi[kz[c] * L.ps + (long long)jy[b] * L.px + ix[a]];
}

template <bool VAR>
__device__ __forceinline__ double node_Lphi_g(const Lev& L, const double* __restrict__ phi, int i, int j, int kl, double& s0)
{
    double P[3][3][3];
    gather_phi_g(L, phi, i, j, kl, P);
    if (VAR) {
        double S[2][2][2];
        gather_sigma_g<true>(L, i, j, kl, S);
        return stencil27(L, S, P, s0);
    }
    return stencil27_c(L, L.csig, P, s0);
}

// ------------------------------------------------------------------------------------------
// block reductions (warp shuffles + one smem hop)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// all threads of the block must call; result valid in every thread
template <bool MAX>
__device__ __forceinline__ double block_reduce(double v, double* sh /* >= 33 doubles */)
{
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = MAX ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        double x = lane < nw ? sh[lane] : (MAX ? 0.0 : 0.0);
        x = MAX ? warp_max(x) : warp_sum(x);
        if (lane == 0) sh[32] = x;
    }
    __syncthreads();
    return sh[32];
}

// ------------------------------------------------------------------------------------------
// K5: full-weighting restriction (A.5).  One thread per coarse node.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_restrict(const Lev F, const Lev C, const double* __restrict__ fine,
                                                  double* __restrict__ crse)
{
    const int i = blockIdx.x * 64 + (threadIdx.x & 63);
    const int j = blockIdx.y * 4 + (threadIdx.x >> 6);
    const int kl = blockIdx.z;  // coarse local plane
    pdl_trigger();
    pdl_wait();
    if (i >= C.nn[0] || j >= C.nn[1]) return;
    long long id = kl * C.ps + (long long)j * C.px + i;
    if (node_masked(C, i, j, kl + C.k0)) { crse[id] = 0.0; return; }
    const int fk = 2 * (kl + C.k0) - F.k0;  // fine local plane
    double s = 0.0;
#pragma unroll
    for (int c = -1; c <= 1; ++c) {
        const double* pl = fine + zplane(F, fk + c) * F.ps;
#pragma unroll
        for (int b = -1; b <= 1; ++b) {
            const double* row = pl + (long long)nmap(2 * j + b, F.n[1], F.per[1]) * F.px;
#pragma unroll
            for (int a = -1; a <= 1; ++a) {
                double w = (a ? 1.0 : 2.0) * (b ? 1.0 : 2.0) * (c ? 1.0 : 2.0);
                s += w * row[nmap(2 * i + a, F.n[0], F.per[0])];
            }
        }
    }
    crse[id] = s * (1.0 / 64.0);
}

// ------------------------------------------------------------------------------------------
// K6, node-by-node form of the sigma-weighted trilinear prolongation (A.6) -- used by the coarse-tail kernel
// (np_tail.cuh); the tiled kernel of the big levels is k_interp_tile (np_smooth.cuh).
// ------------------------------------------------------------------------------------------
template <bool VAR>
struct Interp {
    const Lev& F;
    const Lev& C;
    const double* crse;
    int fk0;  // F.k0
    __device__ __forceinline__ double sg(int i, int j, int kl) const
    {
        if (!VAR) return 1.0;
        return F.sigma[czplane(F, kl) * F.cps + (long long)cmap(j, F.n[1], F.per[1]) * F.cpx + cmap(i, F.n[0], F.per[0])];
    }
    // coarse value at coarse global index (ic, jc, kcg)
    __device__ __forceinline__ double cr(int ic, int jc, int kcg) const
    {
        return crse[zplane(C, kcg - C.k0) * C.ps + (long long)nmap(jc, C.n[1], C.per[1]) * C.px + nmap(ic, C.n[0], C.per[0])];
    }
    __device__ __forceinline__ double qx(int i, int j, int k, int side) const
    {
        int ii = i - 1 + side;
        return sg(ii, j - 1, k - 1) + sg(ii, j, k - 1) + sg(ii, j - 1, k) + sg(ii, j, k);
    }
    __device__ __forceinline__ double qy(int i, int j, int k, int side) const
    {
        int jj = j - 1 + side;
        return sg(i - 1, jj, k - 1) + sg(i, jj, k - 1) + sg(i - 1, jj, k) + sg(i, jj, k);
    }
    __device__ __forceinline__ double qz(int i, int j, int k, int side) const
    {
        int kk = k - 1 + side;
        return sg(i - 1, j - 1, kk) + sg(i, j - 1, kk) + sg(i - 1, j, kk) + sg(i, j, kk);
    }
    // (i,j,k): fine node, k local; (ic,jc,kc): coarse node below it, kc global
    __device__ __forceinline__ double line_x(int i, int j, int k, int ic, int jc, int kc) const
    {
        double w1 = qx(i, j, k, 0), w2 = qx(i, j, k, 1);
        return (w1 * cr(ic, jc, kc) + w2 * cr(ic + 1, jc, kc)) / (w1 + w2);
    }
    __device__ __forceinline__ double line_y(int i, int j, int k, int ic, int jc, int kc) const
    {
        double w1 = qy(i, j, k, 0), w2 = qy(i, j, k, 1);
        return (w1 * cr(ic, jc, kc) + w2 * cr(ic, jc + 1, kc)) / (w1 + w2);
    }
    __device__ __forceinline__ double line_z(int i, int j, int k, int ic, int jc, int kc) const
    {
        double w1 = qz(i, j, k, 0), w2 = qz(i, j, k, 1);
        return (w1 * cr(ic, jc, kc) + w2 * cr(ic, jc, kc + 1)) / (w1 + w2);
    }
    __device__ double face_xy(int i, int j, int k, int ic, int jc, int kc) const
    {
        double w1 = qx(i, j, k, 0), w2 = qx(i, j, k, 1), w3 = qy(i, j, k, 0), w4 = qy(i, j, k, 1);
        return (w1 * line_y(i - 1, j, k, ic, jc, kc) + w2 * line_y(i + 1, j, k, ic + 1, jc, kc) +
                w3 * line_x(i, j - 1, k, ic, jc, kc) + w4 * line_x(i, j + 1, k, ic, jc + 1, kc)) / (w1 + w2 + w3 + w4);
    }
    __device__ double face_xz(int i, int j, int k, int ic, int jc, int kc) const
    {
        double w1 = qx(i, j, k, 0), w2 = qx(i, j, k, 1), w3 = qz(i, j, k, 0), w4 = qz(i, j, k, 1);
        return (w1 * line_z(i - 1, j, k, ic, jc, kc) + w2 * line_z(i + 1, j, k, ic + 1, jc, kc) +
                w3 * line_x(i, j, k - 1, ic, jc, kc) + w4 * line_x(i, j, k + 1, ic, jc, kc + 1)) / (w1 + w2 + w3 + w4);
    }
    __device__ double face_yz(int i, int j, int k, int ic, int jc, int kc) const
    {
        double w1 = qy(i, j, k, 0), w2 = qy(i, j, k, 1), w3 = qz(i, j, k, 0), w4 = qz(i, j, k, 1);
        return (w1 * line_z(i, j - 1, k, ic, jc, kc) + w2 * line_z(i, j + 1, k, ic, jc + 1, kc) +
                w3 * line_y(i, j, k - 1, ic, jc, kc) + w4 * line_y(i, j, k + 1, ic, jc, kc + 1)) / (w1 + w2 + w3 + w4);
    }
    // the interpolant at fine node (i, j, k local plane / kg global plane)
    __device__ double value(int i, int j, int k, int kg) const
    {
        const int ic = i >> 1, jc = j >> 1, kc = kg >> 1;
        const int io = i & 1, jo = j & 1, ko = kg & 1;
        if (io && jo && ko) {
            double w1 = qx(i, j, k, 0), w2 = qx(i, j, k, 1), w3 = qy(i, j, k, 0), w4 = qy(i, j, k, 1), w5 = qz(i, j, k, 0),
                   w6 = qz(i, j, k, 1);
            return (w1 * face_yz(i - 1, j, k, ic, jc, kc) + w2 * face_yz(i + 1, j, k, ic + 1, jc, kc) +
                    w3 * face_xz(i, j - 1, k, ic, jc, kc) + w4 * face_xz(i, j + 1, k, ic, jc + 1, kc) +
                    w5 * face_xy(i, j, k - 1, ic, jc, kc) + w6 * face_xy(i, j, k + 1, ic, jc, kc + 1)) /
                   (w1 + w2 + w3 + w4 + w5 + w6);
        }
        if (jo && ko) return face_yz(i, j, k, ic, jc, kc);
        if (io && ko) return face_xz(i, j, k, ic, jc, kc);
        if (io && jo) return face_xy(i, j, k, ic, jc, kc);
        if (io) return line_x(i, j, k, ic, jc, kc);
        if (jo) return line_y(i, j, k, ic, jc, kc);
        if (ko) return line_z(i, j, k, ic, jc, kc);
        return cr(ic, jc, kc);
    }
};

// ------------------------------------------------------------------------------------------
// K9: sigma on the next-coarser level = arithmetic mean of the 8 children (A.8)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_coarsen_sigma(const Lev F, const Lev C, const double* __restrict__ fs,
                                                       double* __restrict__ cs)
{
    const int i = blockIdx.x * 64 + (threadIdx.x & 63);
    const int j = blockIdx.y * 4 + (threadIdx.x >> 6);
    const int k = blockIdx.z;  // coarse local cell plane
    if (i >= C.n[0] || j >= C.n[1]) return;
    const int fk = 2 * (k + C.ck0) - F.ck0;
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const double2 v = *reinterpret_cast<const double2*>(fs + (fk + c) * F.cps + (long long)(2 * j + b) * F.cpx + 2 * i);
            s += v.x + v.y;
        }
    cs[k * C.cps + (long long)j * C.cpx + i] = 0.125 * s;
}

// ------------------------------------------------------------------------------------------
// caller-side arrays (amrex::Array4 layout)
// ------------------------------------------------------------------------------------------
struct Fab {
    double* p;
    int lo[3];
    int nx, ny, nz;   // extents of the grown box
    long long cstride;
    __host__ __device__ __forceinline__ long long idx(int i, int j, int k, int c) const
    {
        return (i - lo[0]) + (long long)nx * ((j - lo[1]) + (long long)ny * (k - lo[2])) + c * cstride;
    }
    __host__ __device__ __forceinline__ bool has(int i, int j, int k) const
    {
        return i >= lo[0] && i < lo[0] + nx && j >= lo[1] && j < lo[1] + ny && k >= lo[2] && k < lo[2] + nz;
    }
};

// K1: u += s*gp/rho (when !incremental), u -= u_old (when sub_old), sigma = s/rho.
// Also zeroes nothing: ghost handling is done by k_set_vel_ghosts.  One thread per valid cell.
// Traffic: 24 R + 24 R + 8 R + 24 W + 8 W = 88 B/cell (variable density, non-incremental).
__global__ void __launch_bounds__(256) k_pre_add_sigma(const Lev L, Fab vel, Fab gp, Fab rho, Fab velo, double s,
                                                       double ro_0, int add_gp, int sub_old, double* __restrict__ sigma)
{
    const int i = blockIdx.x * 64 + (threadIdx.x & 63);
    const int j = blockIdx.y * 4 + (threadIdx.x >> 6);
    const int kl = blockIdx.z;
    if (i >= L.n[0] || j >= L.n[1]) return;
    const int kg = kl + L.ck0;
    double r = rho.p ? rho.p[rho.idx(i, j, kg, 0)] : ro_0;
    double soverrho = s / r;
    if (sigma) sigma[kl * L.cps + (long long)j * L.cpx + i] = soverrho;
    if (add_gp || sub_old) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            long long vi = vel.idx(i, j, kg, c);
            double u = vel.p[vi];
            if (add_gp) u += gp.p[gp.idx(i, j, kg, c)] * soverrho;
            if (sub_old) u -= velo.p[velo.idx(i, j, kg, c)];
            vel.p[vi] = u;
        }
    }
}

// copies sigma from the caller's box into the level array
__global__ void __launch_bounds__(256) k_copy_sigma(const Lev L, Fab sig, double* __restrict__ sigma)
{
    const int i = blockIdx.x * 64 + (threadIdx.x & 63);
    const int j = blockIdx.y * 4 + (threadIdx.x >> 6);
    const int kl = blockIdx.z;
    if (i >= L.n[0] || j >= L.n[1]) return;
    sigma[kl * L.cps + (long long)j * L.cpx + i] = sig.p[sig.idx(i, j, kl + L.ck0, 0)];
}

// vel.setBndry(0.0) (:137) followed by the inflow fill of the first ghost layer (:138-163).
// One thread per cell of the grown box; only ghost cells are touched.
__global__ void __launch_bounds__(256) k_set_vel_ghosts(const Lev L, Fab vel, Fab inflow, int set_inflow)
{
    const long long total = (long long)vel.nx * vel.ny * vel.nz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int i = (int)(t % vel.nx) + vel.lo[0];
        int j = (int)((t / vel.nx) % vel.ny) + vel.lo[1];
        int k = (int)(t / ((long long)vel.nx * vel.ny)) + vel.lo[2];
        bool out = i < 0 || i >= L.n[0] || j < 0 || j >= L.n[1] || k < 0 || k >= L.n[2];
        if (!out) continue;
        bool infl = false;
        if (set_inflow && inflow.p && i >= -1 && i <= L.n[0] && j >= -1 && j <= L.n[1] && k >= -1 && k <= L.n[2]) {
            infl = (i < 0 && L.rlo[0] == 2) || (i >= L.n[0] && L.rhi[0] == 2) || (j < 0 && L.rlo[1] == 2) ||
                   (j >= L.n[1] && L.rhi[1] == 2) || (k < 0 && L.rlo[2] == 2) || (k >= L.n[2] && L.rhi[2] == 2);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) vel.p[vel.idx(i, j, k, c)] = infl ? inflow.p[inflow.idx(i, j, k, c)] : 0.0;
    }
}

// IncfloVelFill (src/prob/prob_bc.H:8-351) evaluated on the device: fills the FIRST ghost layer of the velocity
// (:138-163 of incflo_apply_nodal_projection.cpp call it with nghost = 1) from the face's boundary velocity
// bcv_vel and the probtype-specific profile of the normal component.  The face blocks are applied in the
// reference's order (x-lo, x-hi, y-lo, y-hi, z-lo, z-hi), so a ghost cell outside the domain in two directions
// ends up with the later face's value, as there.
// bcv[o][c]: o = amrex::Orientation index (dir + 3 * side), c = velocity component (m_bc_velocity).
// face[o]: how ApplyNodalProjection's inflow_bcr sees the face (:141-153):
//   FACE_MASS_INFLOW         BCType::ext_dir            -> boundary value
//   FACE_DIRECTION_DEPENDENT BCType::direction_dependent -> boundary value where the profile points into the domain,
//                            else a copy of the first interior cell (prob_bc.H:93-109 and the five sibling blocks;
//                            the z-lo block tests norm_vel <= 0 for inflow, :300-301 -- kept as written there)
//   anything else            untouched (BCRec default)
// and the "special case" probtypes 1101 (x faces, :86-92, :140-146) / 1102 (y-hi face, :243-251), whose blocks do not
// look at the BCRec at all.
enum { FACE_PLAIN = 0, FACE_MASS_INFLOW = 1, FACE_DIRECTION_DEPENDENT = 2, FACE_MIXED = 3 };
struct InflowProfile {
    int probtype;
    double time;
    double bcv[6][3];
    int face[6];
};
__device__ __forceinline__ double parab6(int idx, int n) { const double s = (idx + 0.5) / n; return 6.0 * s * (1.0 - s); }
// the profile of the normal velocity on face (dir, side) at ghost cell (i,j,k)
__device__ __forceinline__ double inflow_norm_vel(const InflowProfile& P, const Lev& L, int dir, int side, int i, int j, int k)
{
    double norm_vel = P.bcv[dir + 3 * side][dir];
    const int pt = P.probtype;
    if (dir == 0 && side == 0) {          // prob_bc.H:57-84
        if (pt == 42) norm_vel = P.time;
        else if (pt == 31) norm_vel = parab6(j, L.n[1]);
        else if (pt == 43) norm_vel = parab6(j, L.n[1]) - 1.0;
        else if (pt == 311) norm_vel = parab6(k, L.n[2]);
        else if (pt == 41) norm_vel = 0.5 * ((k + 0.5) / L.n[2]);
    } else if (dir == 0 && side == 1) {   // :129-138
        if (pt == 42) norm_vel = P.time;
        else if (pt == 43) norm_vel = parab6(j, L.n[1]) - 1.0;
    } else if (dir == 1 && side == 0) {   // :190-202
        if (pt == 32) norm_vel *= parab6(k, L.n[2]);
        if (pt == 322) norm_vel *= parab6(i, L.n[0]);
    } else if (dir == 1 && side == 1) {   // :241-246
        if (pt == 16) { const double x = (i + 0.5) / L.n[0]; norm_vel = 16.0 * (x * x * x * x - 2.0 * x * x * x + x * x); }
    } else if (dir == 2 && side == 0) {   // :298-308
        if (pt == 33) norm_vel *= parab6(i, L.n[0]);
        else if (pt == 333) norm_vel *= parab6(j, L.n[1]);
    }
    return norm_vel;
}
// does the profile value count as inflow on a direction_dependent face?  (the sign tests of the six blocks)
__device__ __forceinline__ bool dd_is_inflow(int dir, int side, double norm_vel)
{
    if (side == 0 && dir != 2) return norm_vel >= 0.0;   // x-lo :94, y-lo :201
    return norm_vel <= 0.0;                               // x-hi :148, y-hi :254, z-lo :301 (sic), z-hi :333
}
// all components of ghost cell (i,j,k); only components some face block writes are stored
__device__ __forceinline__ void incflo_vel_fill_cell(const InflowProfile& P, const Lev& L, Fab& vel, int i, int j, int k)
{
    const int idx[3] = {i, j, k};
#pragma unroll
    for (int nc = 0; nc < 3; ++nc) {
        bool hit = false;
        double out = 0.0;
#pragma unroll
        for (int dir = 0; dir < 3; ++dir)
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                const bool outside = side == 0 ? idx[dir] < 0 : idx[dir] >= L.n[dir];
                if (!outside || L.per[dir]) continue;
                const int o = dir + 3 * side;
                const int pt = P.probtype;
                if (pt == 1101 && dir == 0) {            // :86-92 / :140-146: all the logic of the x faces is here
                    const int half = L.n[1] / 2;
                    if (side == 0 && j > half) { out = P.bcv[0][nc]; hit = true; }
                    if (side == 1 && j <= half) { out = -P.bcv[3][nc]; hit = true; }
                    continue;
                }
                if (pt == 1102 && dir == 1 && side == 1) {   // :243-251, and :251 skips the generic part
                    if (k <= L.n[2] / 2) { out = -P.bcv[4][nc]; hit = true; }
                    continue;
                }
                const int ft = P.face[o];
                if (ft != FACE_MASS_INFLOW && ft != FACE_DIRECTION_DEPENDENT) continue;
                const double norm_vel = inflow_norm_vel(P, L, dir, side, i, j, k);
                if (ft == FACE_MASS_INFLOW || dd_is_inflow(dir, side, norm_vel)) {
                    out = nc == dir ? norm_vel : P.bcv[o][nc];   // normal component: the profile; tangential: bcv_vel
                } else {   // the flow leaves: first interior cell (a ghost cell of another face for edge / corner cells)
                    int q[3] = {i, j, k};
                    q[dir] += side == 0 ? 1 : -1;
                    out = vel.has(q[0], q[1], q[2]) ? vel.p[vel.idx(q[0], q[1], q[2], nc)] : 0.0;
                }
                hit = true;
            }
        if (hit) vel.p[vel.idx(i, j, k, nc)] = out;
    }
}
// one thread per cell of the domain grown by one; only ghost cells are written
__global__ void __launch_bounds__(256) k_incflo_vel_fill(const Lev L, Fab vel, const InflowProfile P)
{
    const int nx = L.n[0] + 2, ny = L.n[1] + 2;
    const long long total = (long long)nx * ny * (L.cnzl + 2);
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % nx) - 1, j = (int)((t / nx) % ny) - 1, k = (int)(t / ((long long)nx * ny)) - 1 + L.ck0;
        const bool out = i < 0 || i >= L.n[0] || j < 0 || j >= L.n[1] || k < 0 || k >= L.n[2];
        if (!out || !vel.has(i, j, k)) continue;
        incflo_vel_fill_cell(P, L, vel, i, j, k);
    }
}

// HydroUtils::enforceInOutSolvability (AMReX-Hydro, un-vendored; call site incflo_apply_nodal_projection.cpp:166-179):
// over the direction_dependent faces, influx = sum |u_n| dS over the boundary cells where the normal ghost velocity
// points into the domain, outflux = the same over those where it points out; the outflow values are then scaled by
// influx / outflux so that the net flux vanishes (the nodal solve has no Dirichlet node to absorb a net flux).
// face cells: the first ghost layer over the face's own extent (edge / corner ghost cells excluded).
// Pass 1: per-block partial sums {influx, outflux} (fixed order -> deterministic), this rank's cell planes only.
__global__ void __launch_bounds__(256) k_inout_flux(const Lev L, Fab vel, const InflowProfile P, double* __restrict__ partial)
{
    __shared__ double sh[34];
    double fin = 0.0, fout = 0.0;
    const int klo = L.ck0, khi = L.ck0 + L.cnzl;   // owned cell planes [klo, khi)
    for (int o = 0; o < 6; ++o) {
        if (P.face[o] != FACE_DIRECTION_DEPENDENT) continue;
        const int dir = o % 3, side = o / 3;
        const int d1 = (dir + 1) % 3, d2 = (dir + 2) % 3;
        const double ds = 1.0 / (L.dxinv[d1] * L.dxinv[d2]);
        const long long total = (long long)L.n[d1] * L.n[d2];
        for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
            int q[3];
            q[dir] = side == 0 ? -1 : L.n[dir];
            q[d1] = (int)(t % L.n[d1]); q[d2] = (int)(t / L.n[d1]);
            if (dir == 2) { if (side == 0 ? klo != 0 : khi != L.n[2]) continue; }   // the rank at that end of the domain
            else if (q[2] < klo || q[2] >= khi) continue;
            if (!vel.has(q[0], q[1], q[2])) continue;
            const double v = vel.p[vel.idx(q[0], q[1], q[2], dir)];
            const bool in = side == 0 ? v >= 0.0 : v <= 0.0;
            if (in) fin += fabs(v) * ds; else fout += fabs(v) * ds;
        }
    }
    fin = block_reduce<false>(fin, sh);
    fout = block_reduce<false>(fout, sh);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = fin; partial[2 * blockIdx.x + 1] = fout; }
}
// Pass 2: outflow cells *= sums[0] / sums[1] (both above small_vel, else nothing to do / the host reports the error)
__global__ void __launch_bounds__(256) k_inout_correct(const Lev L, Fab vel, const InflowProfile P, const double* __restrict__ sums, double small_vel)
{
    const double influx = sums[0], outflux = sums[1];
    if (!(influx > small_vel && outflux > small_vel)) return;
    const double alpha = influx / outflux;
    const int klo = L.ck0, khi = L.ck0 + L.cnzl;
    for (int o = 0; o < 6; ++o) {
        if (P.face[o] != FACE_DIRECTION_DEPENDENT) continue;
        const int dir = o % 3, side = o / 3;
        const int d1 = (dir + 1) % 3, d2 = (dir + 2) % 3;
        const long long total = (long long)L.n[d1] * L.n[d2];
        for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
            int q[3];
            q[dir] = side == 0 ? -1 : L.n[dir];
            q[d1] = (int)(t % L.n[d1]); q[d2] = (int)(t / L.n[d1]);
            if (dir == 2) { if (side == 0 ? klo != 0 : khi != L.n[2]) continue; }
            else if (q[2] < klo || q[2] >= khi) continue;
            if (!vel.has(q[0], q[1], q[2])) continue;
            const long long id = vel.idx(q[0], q[1], q[2], dir);
            const double v = vel.p[id];
            const bool in = side == 0 ? v >= 0.0 : v <= 0.0;
            if (!in) vel.p[id] = v * alpha;
        }
    }
}

// MLNodeLaplacian::setOversetMask argument check: count the nodes of the caller's mask (1 = solve, 0 = Dirichlet) that
// differ from the mixed-face mask in effect.  Lm: the level-0 descriptor with the domain's own Dirichlet faces cleared
// (make_nodalBC_mask leaves those at 1).  One thread per node of the box.
__global__ void __launch_bounds__(256) k_check_overset(const Lev Lm, const int* __restrict__ mask, int lox, int loy, int loz, int nx, int ny,
                                                       int nz, int* __restrict__ mismatches)
{
    const long long total = (long long)nx * ny * nz;
    int bad = 0;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % nx) + lox, j = (int)((t / nx) % ny) + loy, k = (int)(t / ((long long)nx * ny)) + loz;
        if (i < 0 || i > Lm.n[0] || j < 0 || j > Lm.n[1] || k < 0 || k > Lm.n[2]) continue;
        const bool dir = node_masked(Lm, nmap(i, Lm.n[0], Lm.per[0]), nmap(j, Lm.n[1], Lm.per[1]), nmap(k, Lm.n[2], Lm.per[2]));
        if ((mask[t] == 0) != dir) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// K2: rhs = D u with the Neumann/inflow treatment of A.2.  One thread per owned node.
// Traffic 24 B/cell R + 8 B/node W.
__global__ void __launch_bounds__(256) k_divu(const Lev L, Fab vel, double* __restrict__ rhs)
{
    const int i = blockIdx.x * 64 + (threadIdx.x & 63);
    const int j = blockIdx.y * 4 + (threadIdx.x >> 6);
    const int kl = blockIdx.z;
    if (i >= L.nn[0] || j >= L.nn[1]) return;
    const int kg = kl + L.k0;
    long long id = kl * L.ps + (long long)j * L.px + i;
    if (node_masked(L, i, j, kg)) { rhs[id] = 0.0; return; }
    // interior node (the 8 cells around it are domain cells, held by the caller's box): fixed strides, no index maps
    if (i > 0 && i < L.n[0] && j > 0 && j < L.n[1] && kg > 0 && kg < L.n[2] && (!L.dist || (kg > L.ck0 && kg < L.ck0 + L.cnzl))) {
        const double* p0 = vel.p + vel.idx(i - 1, j - 1, kg - 1, 0);
        const long long sy = vel.nx, sz = (long long)vel.nx * vel.ny;
        double dux = 0, dvy = 0, dwz = 0;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const double* q = p0 + c * sz + b * sy;
                dux += q[1] - q[0];
            }
        const double* pv = p0 + vel.cstride;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int a = 0; a < 2; ++a) dvy += pv[c * sz + sy + a] - pv[c * sz + a];
        const double* pw = p0 + 2 * vel.cstride;
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int a = 0; a < 2; ++a) dwz += pw[sz + b * sy + a] - pw[b * sy + a];
        rhs[id] = 0.25 * (L.dxinv[0] * dux + L.dxinv[1] * dvy + L.dxinv[2] * dwz);
        return;
    }
    double zx[2] = {1.0, 1.0}, zy[2] = {1.0, 1.0}, zz[2] = {1.0, 1.0}, scale = 1.0;
    if (!L.per[0]) { if (i == 0 && L.rlo[0]) { zx[0] = 0; scale *= 2; } if (i == L.n[0] && L.rhi[0]) { zx[1] = 0; scale *= 2; } }
    if (!L.per[1]) { if (j == 0 && L.rlo[1]) { zy[0] = 0; scale *= 2; } if (j == L.n[1] && L.rhi[1]) { zy[1] = 0; scale *= 2; } }
    if (!L.per[2]) { if (kg == 0 && L.rlo[2]) { zz[0] = 0; scale *= 2; } if (kg == L.n[2] && L.rhi[2]) { zz[1] = 0; scale *= 2; } }
    // cell coordinates of the 2x2x2 block around the node; periodic ghosts wrap (FillBoundary)
    int ci[2] = {i - 1, i}, cj[2] = {j - 1, j}, ck[2] = {kg - 1, kg};
    if (L.per[0]) { ci[0] = cmap(ci[0], L.n[0], 1); ci[1] = cmap(ci[1], L.n[0], 1); }
    if (L.per[1]) { cj[0] = cmap(cj[0], L.n[1], 1); cj[1] = cmap(cj[1], L.n[1], 1); }
    if (L.per[2] && !L.dist) { ck[0] = cmap(ck[0], L.n[2], 1); ck[1] = cmap(ck[1], L.n[2], 1); }
    double u[2][2][2], v[2][2][2], w[2][2][2];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                bool ok = vel.has(ci[a], cj[b], ck[c]);
                long long q = ok ? vel.idx(ci[a], cj[b], ck[c], 0) : 0;
                u[c][b][a] = ok ? vel.p[q] : 0.0;
                v[c][b][a] = ok ? vel.p[q + vel.cstride] : 0.0;
                w[c][b][a] = ok ? vel.p[q + 2 * vel.cstride] : 0.0;
            }
    double dux = 0, dvy = 0, dwz = 0;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            dux += (u[c][b][1] - u[c][b][0]) * zy[b] * zz[c];
            dvy += (v[c][1][b] - v[c][0][b]) * zx[b] * zz[c];
            dwz += (w[1][c][b] - w[0][c][b]) * zx[b] * zy[c];
        }
    rhs[id] = scale * 0.25 * (L.dxinv[0] * dux + L.dxinv[1] * dvy + L.dxinv[2] * dwz);
}

// K7: fused final update, one thread per valid cell:
//   g = G phi;  vel -= sigma g;  (vel += vel_old);  gphi (=|+=) g
// Traffic: 8 R phi + 8 R sigma + 24 R + 24 W vel + 24 W gp = 88 B/cell (+24 R when accumulating).
__global__ void __launch_bounds__(256) k_mknewu(const Lev L, const double* __restrict__ phi, Fab vel, Fab velo, int add_old,
                                                Fab gphi, int accumulate)
{
    const int i = blockIdx.x * 64 + (threadIdx.x & 63);
    const int j = blockIdx.y * 4 + (threadIdx.x >> 6);
    const int kl = blockIdx.z;  // local cell plane
    if (i >= L.n[0] || j >= L.n[1]) return;
    const int kg = kl + L.ck0;
    const int nl = kg - L.k0;  // local node plane of the cell's lower face
    int ix[2] = {i, nmap(i + 1, L.n[0], L.per[0])}, jy[2] = {j, nmap(j + 1, L.n[1], L.per[1])};
    int kz[2] = {nl, zplane(L, nl + 1)};
    double P[2][2][2];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int a = 0; a < 2; ++a) P[c][b][a] = phi[kz[c] * L.ps + (long long)jy[b] * L.px + ix[a]];
    double g[3];
    g[0] = 0.25 * L.dxinv[0] * ((P[0][0][1] - P[0][0][0]) + (P[0][1][1] - P[0][1][0]) + (P[1][0][1] - P[1][0][0]) + (P[1][1][1] - P[1][1][0]));
    g[1] = 0.25 * L.dxinv[1] * ((P[0][1][0] - P[0][0][0]) + (P[0][1][1] - P[0][0][1]) + (P[1][1][0] - P[1][0][0]) + (P[1][1][1] - P[1][0][1]));
    g[2] = 0.25 * L.dxinv[2] * ((P[1][0][0] - P[0][0][0]) + (P[1][0][1] - P[0][0][1]) + (P[1][1][0] - P[0][1][0]) + (P[1][1][1] - P[0][1][1]));
    double sg = L.sigma ? L.sigma[kl * L.cps + (long long)j * L.cpx + i] : L.csig;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (vel.p) {
            long long vi = vel.idx(i, j, kg, c);
            double u = vel.p[vi] - sg * g[c];
            if (add_old) u += velo.p[velo.idx(i, j, kg, c)];
            vel.p[vi] = u;
        }
        if (gphi.p) {
            long long gi = gphi.idx(i, j, kg, c);
            gphi.p[gi] = accumulate ? gphi.p[gi] + g[c] : g[c];
        }
    }
}

// y += x on the valid cells of two caller arrays (vel += vel_old after the composite average-down)
__global__ void __launch_bounds__(256) k_add_cells(const Lev L, Fab y, Fab x, int ncomp)
{
    const int i = blockIdx.x * 64 + (threadIdx.x & 63);
    const int j = blockIdx.y * 4 + (threadIdx.x >> 6);
    const int kg = blockIdx.z + L.ck0;
    if (i >= L.n[0] || j >= L.n[1]) return;
    for (int c = 0; c < ncomp; ++c) y.p[y.idx(i, j, kg, c)] += x.p[x.idx(i, j, kg, c)];
}

// copy-out of phi into the caller's nodal box (p_nd (=|+=) phi, :235-253); duplicates the
// periodic image nodes that the unique-node storage does not hold.
__global__ void __launch_bounds__(256) k_copy_phi(const Lev L, const double* __restrict__ phi, Fab out, int accumulate)
{
    const int i = blockIdx.x * 64 + (threadIdx.x & 63) + out.lo[0];
    const int j = blockIdx.y * 4 + (threadIdx.x >> 6) + out.lo[1];
    const int k = blockIdx.z + out.lo[2];
    if (i >= out.lo[0] + out.nx || j >= out.lo[1] + out.ny) return;
    if (i < 0 || i > L.n[0] || j < 0 || j > L.n[1] || k < 0 || k > L.n[2]) return;
    if (L.dist && (k - L.k0 < -1 || k - L.k0 > L.nzl)) return;   // beyond this slab's planes and ghost slots
    int kl = zplane(L, k - L.k0);
    double v = phi[kl * L.ps + (long long)nmap(j, L.n[1], L.per[1]) * L.px + nmap(i, L.n[0], L.per[0])];
    long long o = out.idx(i, j, k, 0);
    out.p[o] = accumulate ? out.p[o] + v : v;
}

// ------------------------------------------------------------------------------------------
// small vector kernels over the owned nodes of a level
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_axpy(const Lev L, double* __restrict__ y, const double* __restrict__ x, double a)
{
    const int i = blockIdx.x * 64 + (threadIdx.x & 63);
    const int j = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (i >= L.nn[0] || j >= L.nn[1]) return;
    long long id = blockIdx.z * L.ps + (long long)j * L.px + i;
    y[id] += a * x[id];
}

// partial[block] = sum w*x  and  sum w   (weighted mean for the solvability offset, A.8)
__global__ void __launch_bounds__(256) k_wsum_partial(const Lev L, const double* __restrict__ x, double* __restrict__ partial)
{
    __shared__ double sh[34];
    const int i = blockIdx.x * 64 + (threadIdx.x & 63);
    const int j = blockIdx.y * 4 + (threadIdx.x >> 6);
    const int kl = blockIdx.z;
    double w = 0.0, v = 0.0;
    if (i < L.nn[0] && j < L.nn[1]) {
        w = node_weight(L, i, j, kl + L.k0);
        v = w * x[kl * L.ps + (long long)j * L.px + i];
    }
    v = block_reduce<false>(v, sh);
    w = block_reduce<false>(w, sh);
    if (threadIdx.x == 0) {
        long long b = (blockIdx.z * gridDim.y + blockIdx.y) * (long long)gridDim.x + blockIdx.x;
        partial[2 * b] = v; partial[2 * b + 1] = w;
    }
}
// out[0] = sum partial[2b], out[1] = sum partial[2b+1] in a fixed order (deterministic)
__global__ void __launch_bounds__(1024) k_sum2_final(const double* __restrict__ partial, long long nb, double* __restrict__ out)
{
    __shared__ double sh[34];
    double a = 0, b = 0;
    for (long long t = threadIdx.x; t < nb; t += blockDim.x) { a += partial[2 * t]; b += partial[2 * t + 1]; }
    a = block_reduce<false>(a, sh);
    b = block_reduce<false>(b, sh);
    if (threadIdx.x == 0) { out[0] = a; out[1] = b; }
}
__global__ void __launch_bounds__(1024) k_max_final(const double* __restrict__ partial, long long nb, double* __restrict__ out)
{
    __shared__ double sh[34];
    double a = 0;
    for (long long t = threadIdx.x; t < nb; t += blockDim.x) a = fmax(a, partial[t]);
    a = block_reduce<true>(a, sh);
    if (threadIdx.x == 0) out[0] = a;
}
// x -= sums[0]/sums[1] on unmasked owned nodes
// partial (optional): per-block max |x| after the subtraction, laid out like k_norminf_partial's -- the inf-norm in the same pass
__global__ void __launch_bounds__(256) k_sub_mean(const Lev L, double* __restrict__ x, const double* __restrict__ sums,
                                                  double* __restrict__ partial = nullptr)
{
    __shared__ double sh[34];
    const int i = blockIdx.x * 64 + (threadIdx.x & 63);
    const int j = blockIdx.y * 4 + (threadIdx.x >> 6);
    double a = 0.0;
    if (i < L.nn[0] && j < L.nn[1]) {
        const long long id = blockIdx.z * L.ps + (long long)j * L.px + i;
        const double v = x[id] - sums[0] / sums[1];
        x[id] = v;
        a = fabs(v);
    }
    if (partial) {
        a = block_reduce<true>(a, sh);
        if (threadIdx.x == 0) partial[(blockIdx.z * gridDim.y + blockIdx.y) * (long long)gridDim.x + blockIdx.x] = a;
    }
}
__global__ void __launch_bounds__(256) k_norminf_partial(const Lev L, const double* __restrict__ x, double* __restrict__ partial)
{
    __shared__ double sh[34];
    const int i = blockIdx.x * 64 + (threadIdx.x & 63);
    const int j = blockIdx.y * 4 + (threadIdx.x >> 6);
    double a = 0.0;
    if (i < L.nn[0] && j < L.nn[1]) a = fabs(x[blockIdx.z * L.ps + (long long)j * L.px + i]);
    a = block_reduce<true>(a, sh);
    if (threadIdx.x == 0) partial[(blockIdx.z * gridDim.y + blockIdx.y) * (long long)gridDim.x + blockIdx.x] = a;
}
// zero masked (Dirichlet) nodes of x
__global__ void __launch_bounds__(256) k_zero_masked(const Lev L, double* __restrict__ x)
{
    const int i = blockIdx.x * 64 + (threadIdx.x & 63);
    const int j = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (i >= L.nn[0] || j >= L.nn[1]) return;
    if (node_masked(L, i, j, blockIdx.z + L.k0)) x[blockIdx.z * L.ps + (long long)j * L.px + i] = 0.0;
}

// ------------------------------------------------------------------------------------------
// K8: bottom solve on the coarsest level in ONE CTA: BiCGStab (A.10) with a CG retry
// (bottom_solver = bicgcg), all dot products / norms by warp-shuffle block reductions, no host
// round trips.  The coarsest level lives on a single GPU (after agglomeration in the
// multi-GPU case), so there is no collective either.
//   work: 8 vectors of L.ps*nzl doubles (r, rh, p, v, s, t, b, tmp)
//   info[0] = iterations, info[1] = return code (0 ok)
// ------------------------------------------------------------------------------------------
template <bool VAR>
__device__ void bottom_apply(const Lev& L, const double* __restrict__ x, double* __restrict__ y)
{
    const int nxy = L.nn[0] * L.nn[1], ntot = nxy * L.nzl;
    for (int t = threadIdx.x; t < ntot; t += blockDim.x) {
        int kl = t / nxy, r = t - kl * nxy, j = r / L.nn[0], i = r - j * L.nn[0];
        double s0, v = 0.0;
        if (!node_masked(L, i, j, kl + L.k0)) v = node_Lphi_g<VAR>(L, x, i, j, kl, s0);
        y[kl * L.ps + (long long)j * L.px + i] = v;
    }
    __syncthreads();
}

#define BOT_FOR_NODES(body)                                                        \
    for (int t = threadIdx.x; t < ntot; t += blockDim.x) {                         \
        int kl = t / nxy, rr = t - kl * nxy, j = rr / L.nn[0], i = rr - j * L.nn[0]; \
        long long id = kl * L.ps + (long long)j * L.px + i;                        \
        (void)i; (void)j; (void)kl;                                                \
        body                                                                       \
    }

template <bool VAR>
__device__ void bottom_bicgstab_body(const Lev& L, double* __restrict__ x, const double* __restrict__ b, double* __restrict__ work,
                                     int maxiter, double rtol, double atol, int singular, int nsweeps, int bottom_solver,
                                     int* __restrict__ info, double* sh /* >= 33 doubles of shared memory */)
{
    const int nxy = L.nn[0] * L.nn[1], ntot = nxy * L.nzl;
    const long long vs = L.ps * L.nzl;
    double *r = work, *rh = work + vs, *p = work + 2 * vs, *v = work + 3 * vs, *s = work + 4 * vs, *t_ = work + 5 * vs,
           *bb = work + 6 * vs;
    // make the rhs solvable: subtract the weighted mean (A.8)
    {
        double sw = 0, sv = 0;
        BOT_FOR_NODES({ double w = node_weight(L, i, j, kl + L.k0); sw += w; sv += w * b[id]; })
        sv = block_reduce<false>(sv, sh);
        sw = block_reduce<false>(sw, sh);
        double off = singular ? sv / sw : 0.0;
        BOT_FOR_NODES({ bb[id] = node_masked(L, i, j, kl + L.k0) ? 0.0 : b[id] - off; })
        __syncthreads();
    }
    int total_iters = 0, ret = 0;
    if (bottom_solver == 1) ret = 9;  // nodal_proj.bottom_solver = smoother
    for (int attempt = 0; attempt < 2 && bottom_solver != 1; ++attempt) {
        // attempt 0: BiCGStab, attempt 1: CG (bicgcg fallback)
        double rn = 0;
        BOT_FOR_NODES({ x[id] = 0.0; r[id] = bb[id]; rh[id] = bb[id]; rn = fmax(rn, fabs(bb[id])); })
        const double rnorm0 = block_reduce<true>(rn, sh);
        ret = 0;
        if (rnorm0 == 0.0 || rnorm0 < atol) break;
        double rho1 = 0, alpha = 0, omega = 0;
        int it = 1;
        bool done = false;
        for (; it <= maxiter && !done; ++it) {
            if (attempt == 0) {
                double d = 0;
                BOT_FOR_NODES({ d += node_weight(L, i, j, kl + L.k0) * rh[id] * r[id]; })
                const double rho = block_reduce<false>(d, sh);
                if (rho == 0.0) { ret = 1; break; }
                if (it == 1) { BOT_FOR_NODES({ p[id] = r[id]; }) }
                else {
                    const double beta = (rho / rho1) * (alpha / omega);
                    BOT_FOR_NODES({ p[id] = r[id] + beta * (p[id] - omega * v[id]); })
                }
                __syncthreads();
                bottom_apply<VAR>(L, p, v);
                d = 0;
                BOT_FOR_NODES({ d += node_weight(L, i, j, kl + L.k0) * rh[id] * v[id]; })
                const double rhTv = block_reduce<false>(d, sh);
                if (rhTv == 0.0) { ret = 3; break; }
                alpha = rho / rhTv;
                double m = 0;
                BOT_FOR_NODES({ x[id] += alpha * p[id]; double sv_ = r[id] - alpha * v[id]; s[id] = sv_; m = fmax(m, fabs(sv_)); })
                double rnorm = block_reduce<true>(m, sh);
                if (rnorm < rtol * rnorm0 || rnorm < atol) { done = true; break; }
                bottom_apply<VAR>(L, s, t_);
                double d1 = 0, d2 = 0;
                BOT_FOR_NODES({ double w = node_weight(L, i, j, kl + L.k0); d1 += w * t_[id] * t_[id]; d2 += w * t_[id] * s[id]; })
                const double tt = block_reduce<false>(d1, sh);
                const double ts = block_reduce<false>(d2, sh);
                if (tt == 0.0) { ret = 4; break; }
                omega = ts / tt;
                m = 0;
                BOT_FOR_NODES({ x[id] += omega * s[id]; double rv = s[id] - omega * t_[id]; r[id] = rv; m = fmax(m, fabs(rv)); })
                rnorm = block_reduce<true>(m, sh);
                if (rnorm < rtol * rnorm0 || rnorm < atol) { done = true; break; }
                if (omega == 0.0) { ret = 4; break; }
                rho1 = rho;
            } else {
                double d = 0;
                BOT_FOR_NODES({ d += node_weight(L, i, j, kl + L.k0) * r[id] * r[id]; })
                const double rho = block_reduce<false>(d, sh);
                if (rho == 0.0) { ret = 1; break; }
                if (it == 1) { BOT_FOR_NODES({ p[id] = r[id]; }) }
                else { const double beta = rho / rho1; BOT_FOR_NODES({ p[id] = r[id] + beta * p[id]; }) }
                __syncthreads();
                bottom_apply<VAR>(L, p, v);
                d = 0;
                BOT_FOR_NODES({ d += node_weight(L, i, j, kl + L.k0) * p[id] * v[id]; })
                const double pq = block_reduce<false>(d, sh);
                if (pq == 0.0) { ret = 1; break; }
                alpha = rho / pq;
                double m = 0;
                BOT_FOR_NODES({ x[id] += alpha * p[id]; double rv = r[id] - alpha * v[id]; r[id] = rv; m = fmax(m, fabs(rv)); })
                const double rnorm = block_reduce<true>(m, sh);
                if (rnorm < rtol * rnorm0 || rnorm < atol) { done = true; break; }
                rho1 = rho;
            }
        }
        total_iters += (it > maxiter ? maxiter : it);
        if (!done && ret == 0) ret = 8;
        __syncthreads();
        if (ret == 0) break;
    }
    if (ret != 0) {
        // MLMG::actualBottomSolve fallback: cor = 0 then 8 smooth calls (A.9), here 8-colour GS in the CTA
        BOT_FOR_NODES({ x[id] = 0.0; })
        __syncthreads();
        for (int sw = 0; sw < 8 * nsweeps; ++sw)
            for (int color = 0; color < 8; ++color) {
                BOT_FOR_NODES({
                    if (((i & 1) + 2 * (j & 1) + 4 * ((kl + L.k0) & 1)) == color) {
                        if (node_masked(L, i, j, kl + L.k0)) x[id] = 0.0;
                        else { double s0; double Ax = node_Lphi_g<VAR>(L, x, i, j, kl, s0); x[id] += (bb[id] - Ax) / s0; }
                    }
                })
                __syncthreads();
            }
    }
    __syncthreads();
    if (threadIdx.x == 0) { atomicAdd(&info[0], total_iters); info[1] = ret; }
}

template <bool VAR>
__global__ void __launch_bounds__(512) k_bottom_bicgstab(const Lev L, double* __restrict__ x, const double* __restrict__ b,
                                                         double* __restrict__ work, int maxiter, double rtol, double atol,
                                                         int singular, int nsweeps, int bottom_solver, int* __restrict__ info)
{
    __shared__ double sh[34];
    pdl_trigger();
    pdl_wait();
    bottom_bicgstab_body<VAR>(L, x, b, work, maxiter, rtol, atol, singular, nsweeps, bottom_solver, info, sh);
}

// ------------------------------------------------------------------------------------------
// K10: FillBoundary between z slabs over NVLink peer memory (replaces a grouped ncclSend/ncclRecv
// pair per halo; MLNodeLinOp::applyBC's FillBoundary).  One process per GPU; every rank
// maps its two neighbours' arenas with CUDA IPC.  Pull protocol, one kernel per exchange:
//   1. thread 0 of CTA 0 tells both neighbours "everything I launched before
//      exchange e is complete" (st.release.sys of e into their flag words, over NVLink);
//   2. every CTA waits until both neighbours have said the same (ld.acquire.sys on MY flag words,
//      local polling), then copies the neighbours' boundary planes into my ghost plane slots with
//      peer loads;
//   (epoch e = base my[2] + 1 + k, k = kernel argument: see k_epoch_advance).
// RAW: step 2.  WAR (a neighbour overwriting the plane I am still pulling): every overwrite of an
// exchanged array is separated from its last pull by at least one later exchange with the same
// neighbour (ping-pong sweeps; see exchange_planes), whose step 1 is stream-ordered after my pull.
// All state lives in device memory, so the kernel can be replayed from a CUDA graph.
// At a physical (non-periodic) end the "source" is the local reflection / clamp plane and the
// flag pointer is null.
// ------------------------------------------------------------------------------------------
// handshake part of an exchange (steps 1 and 2)
__device__ __forceinline__ void halo_handshake(const HaloFlags& f)
{
    const unsigned long long e = halo_epoch(f);
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) {
            __threadfence_system();
            if (f.lo_flag) st_release_sys(f.lo_flag, e);
            if (f.hi_flag) st_release_sys(f.hi_flag, e);
        }
        if (f.lo_flag) halo_spin(f.my, 0, e);
        if (f.hi_flag) halo_spin(f.my, 1, e);
    }
    __syncthreads();
}
// n2 = doubles per plane / 2 (planes are multiples of 64 B and 1 KiB aligned)
__global__ void __launch_bounds__(256) k_halo_pull(const HaloFlags f, double2* __restrict__ ghost_lo, const double2* __restrict__ src_lo,
                                                   double2* __restrict__ ghost_hi, const double2* __restrict__ src_hi, long long n2)
{
    pdl_trigger();   // the successor may become resident; it still waits for this grid in its own pdl_wait()
    pdl_wait();      // the flag below promises that everything launched before this exchange is complete
    halo_handshake(f);
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n2; i += gridDim.x * 256ll) {
        const double2 a = __ldcv(src_lo + i), b = __ldcv(src_hi + i);   // .cv: never served from a stale line
        ghost_lo[i] = a; ghost_hi[i] = b;
    }
}
// Epochs: an exchange kernel gets its number k since the last advance as an argument (so a captured
// CUDA graph can be replayed) and the base lives in device memory; the base is advanced by the number
// of exchanges issued, as the last node of the V-cycle graph and before a graph is launched.
__global__ void k_epoch_advance(unsigned long long* my, unsigned long long n) { my[2] += n; }

}  // namespace b200np_dev
