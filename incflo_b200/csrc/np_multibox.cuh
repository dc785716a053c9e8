// np_multibox.cuh -- K12: gather / scatter between the caller's multi-box MultiFabs and the solver's slab arrays.
//
// Every reference deck runs with amr.max_grid_size = 16 (test_no_eb_3d/benchmark.rayleigh_taylor:16): a MultiFab is
// many small FArrayBoxes, each with its own ghost frame.  The multigrid wants one array per field and rank, so the
// boxes are gathered into the slab arrays on entry and the results scattered back on exit -- one fused pass per
// field and direction, one CTA column per fab (blockIdx.y), all components in the same thread.
#pragma once
#include "np_kernels.cuh"

namespace b200np_dev {

// one FArrayBox: data pointer (device), allocated box
struct MfFab {
    double* p;
    int lo[3];
    int nx, ny, nz;
    long long cstride;
};

enum { MF_VALID = 0,       // the valid box only (gp, p_nd, sigma, density, velocity_o)
       MF_VALID_BC = 1 };  // gather: valid cells + ghost cells OUTSIDE the domain (the BC ghost layer of vel, inflow_vel);
                           // scatter: every cell of the allocated box -- the slab value where the slab array has one
                           // (valid cells, first ghost layer inside the domain grown by one), 0 elsewhere (setBndry(0))

// dst (one box) <- fabs.  nd = 1 for nodal MultiFabs (valid box = cells' box + 1 node at the high ends, which the
// allocated box already contains), dom[d] = number of cells of the domain.
__global__ void __launch_bounds__(256) k_mf_gather(const MfFab* __restrict__ tab, int ngrow, int ncomp, Fab dst, int mode, int dom0, int dom1,
                                                   int dom2)
{
    const MfFab f = tab[blockIdx.y];
    const long long total = (long long)f.nx * f.ny * f.nz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int li = (int)(t % f.nx), lj = (int)((t / f.nx) % f.ny), lk = (int)(t / ((long long)f.nx * f.ny));
        const int i = li + f.lo[0], j = lj + f.lo[1], k = lk + f.lo[2];
        const bool valid = li >= ngrow && li < f.nx - ngrow && lj >= ngrow && lj < f.ny - ngrow && lk >= ngrow && lk < f.nz - ngrow;
        bool take = valid;
        if (!take && mode == MF_VALID_BC) take = i < 0 || i >= dom0 || j < 0 || j >= dom1 || k < 0 || k >= dom2;
        if (!take || !dst.has(i, j, k)) continue;
        for (int c = 0; c < ncomp; ++c) dst.p[dst.idx(i, j, k, c)] = f.p[t + c * f.cstride];
    }
}

// fabs <- src (one box)
__global__ void __launch_bounds__(256) k_mf_scatter(const MfFab* __restrict__ tab, int ngrow, int ncomp, Fab src, int mode)
{
    const MfFab f = tab[blockIdx.y];
    const long long total = (long long)f.nx * f.ny * f.nz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int li = (int)(t % f.nx), lj = (int)((t / f.nx) % f.ny), lk = (int)(t / ((long long)f.nx * f.ny));
        const int i = li + f.lo[0], j = lj + f.lo[1], k = lk + f.lo[2];
        const bool valid = li >= ngrow && li < f.nx - ngrow && lj >= ngrow && lj < f.ny - ngrow && lk >= ngrow && lk < f.nz - ngrow;
        if (!valid && mode != MF_VALID_BC) continue;
        const bool have = src.has(i, j, k);
        if (!have && valid) continue;   // cannot happen for boxes that passed the host checks
        for (int c = 0; c < ncomp; ++c) f.p[t + c * f.cstride] = have ? src.p[src.idx(i, j, k, c)] : 0.0;
    }
}

}  // namespace b200np_dev
