// np_composite.cuh -- box kernels of the composite (two AMR level, one fine box, ratio 2) nodal
// projection (BASELINE configs[3]; multi-level branch of Hydro::NodalProjector / MLMG::oneIter /
// MLNodeLaplacian::reflux, un-vendored AMReX; restated in oracle/composite.py, pinned by
// tests/golden/composite/*.npz).  Everything heavy (smoother, residual, restriction, divergence,
// interpolation, gradient) is done by the single-level kernels applied to
//   * the coarse level with sigma = 0 / u = 0 in the covered cells  -> sums over uncovered coarse cells,
//   * the fine box taken as a domain with reflecting faces          -> 2^f x sums over the fine cells of the box
//     at a node on f box faces, and 2^f x their full-weighting restriction,
// so the kernels here only combine, mask, inject and average.
#pragma once
#include "np_kernels.cuh"

namespace b200np_dev {

// Per direction the box either has a coarse/fine interface on a side (cf_lo / cf_hi), or touches a wall /
// outflow face of the domain there (the fine level inherits the BC), or spans a periodic direction.
struct CBox {
    int lo[3], hi[3];        // covered coarse cells, inclusive
    int cf_lo[3], cf_hi[3];  // 1: coarse/fine interface on this side
    int span[3];             // 1: the box spans this (periodic) direction
    int nbn[3];              // coarse nodes of the box per direction: hi-lo+2, or n when it spans a periodic direction
};

// number of coarse/fine interface faces the coarse node (i,j,k) lies on; -1 if outside the closed box
__device__ __forceinline__ int box_faces(const CBox& b, int i, int j, int k)
{
    const int id[3] = {i, j, k};
    int f = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (b.span[d]) continue;
        if (id[d] < b.lo[d] || id[d] > b.hi[d] + 1) return -1;
        f += (b.cf_lo[d] && id[d] == b.lo[d]) + (b.cf_hi[d] && id[d] == b.hi[d] + 1);
    }
    return f;
}

// reflux: res0 (sums over uncovered coarse cells) += R / 2^f on the box nodes, R = restriction of the
// reflected fine-side residual; then the solvability offset sums[0]/sums[1] is subtracted everywhere.
__global__ void __launch_bounds__(256) k_comp_combine(const Lev L, const CBox b, double* __restrict__ res0,
                                                      const double* __restrict__ R, const double* __restrict__ sums)
{
    const int i = blockIdx.x * 64 + (threadIdx.x & 63);
    const int j = blockIdx.y * 4 + (threadIdx.x >> 6);
    const int k = blockIdx.z;
    if (i >= L.nn[0] || j >= L.nn[1]) return;
    const long long id = k * L.ps + (long long)j * L.px + i;
    double v = res0[id];
    const int f = box_faces(b, i, j, k);
    if (f >= 0) v += R[id] * (f == 0 ? 1.0 : f == 1 ? 0.5 : f == 2 ? 0.25 : 0.125);
    if (sums) v -= sums[0] / sums[1];
    res0[id] = node_masked(L, i, j, k) ? 0.0 : v;
}

// inf-norm partials over the coarse nodes that are not strictly inside the box
__global__ void __launch_bounds__(256) k_comp_norm_excl(const Lev L, const CBox b, const double* __restrict__ x,
                                                        double* __restrict__ partial)
{
    __shared__ double sh[34];
    const int i = blockIdx.x * 64 + (threadIdx.x & 63);
    const int j = blockIdx.y * 4 + (threadIdx.x >> 6);
    const int k = blockIdx.z;
    double a = 0.0;
    if (i < L.nn[0] && j < L.nn[1] && box_faces(b, i, j, k) != 0) a = fabs(x[k * L.ps + (long long)j * L.px + i]);
    a = block_reduce<true>(a, sh);
    if (threadIdx.x == 0) partial[(blockIdx.z * gridDim.y + blockIdx.y) * (long long)gridDim.x + blockIdx.x] = a;
}

// covered coarse cells of a 3-component cell array <- 0 (their input values never count: the
// projected coarse velocity there is the average of the fine one)
__global__ void __launch_bounds__(256) k_comp_zero_cells(const CBox b, Fab v, int ncomp)
{
    const int i = b.lo[0] + blockIdx.x * 64 + (threadIdx.x & 63);
    const int j = b.lo[1] + blockIdx.y * 4 + (threadIdx.x >> 6);
    const int k = b.lo[2] + blockIdx.z;
    if (i > b.hi[0] || j > b.hi[1]) return;
    for (int c = 0; c < ncomp; ++c) v.p[v.idx(i, j, k, c)] = 0.0;
}

// sigma on the covered coarse cells: s0 <- mean of the 8 fine children (or csig when s1 == nullptr),
// s0z <- s0 outside the box and 0 inside.  One thread per coarse cell of the level.
__global__ void __launch_bounds__(256) k_comp_sigma(const Lev L0, const Lev L1, const CBox b, double* __restrict__ s0,
                                                    double* __restrict__ s0z, const double* __restrict__ s1, double csig)
{
    const int i = blockIdx.x * 64 + (threadIdx.x & 63);
    const int j = blockIdx.y * 4 + (threadIdx.x >> 6);
    const int k = blockIdx.z;
    if (i >= L0.n[0] || j >= L0.n[1]) return;
    const long long id = k * L0.cps + (long long)j * L0.cpx + i;
    const bool in = i >= b.lo[0] && i <= b.hi[0] && j >= b.lo[1] && j <= b.hi[1] && k >= b.lo[2] && k <= b.hi[2];
    if (!in) { s0z[id] = s0 ? s0[id] : csig; return; }
    s0z[id] = 0.0;
    if (s0 && s1) {
        const int fi = 2 * (i - b.lo[0]), fj = 2 * (j - b.lo[1]), fk = 2 * (k - b.lo[2]);
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int bb = 0; bb < 2; ++bb)
#pragma unroll
                for (int a = 0; a < 2; ++a) s += s1[(fk + c) * L1.cps + (long long)(fj + bb) * L1.cpx + fi + a];
        s0[id] = 0.125 * s;
    }
}

// injection of the fine solution onto the box nodes of the coarse one (nodal average_down)
__global__ void __launch_bounds__(256) k_comp_inject(const Lev L0, const Lev L1, const CBox b, double* __restrict__ sol0,
                                                     const double* __restrict__ sol1)
{
    const int i = b.lo[0] + blockIdx.x * 64 + (threadIdx.x & 63);
    const int j = b.lo[1] + blockIdx.y * 4 + (threadIdx.x >> 6);
    const int k = b.lo[2] + blockIdx.z;
    if (i >= b.lo[0] + b.nbn[0] || j >= b.lo[1] + b.nbn[1]) return;
    sol0[k * L0.ps + (long long)j * L0.px + i] =
        sol1[2 * (k - b.lo[2]) * L1.ps + (long long)(2 * (j - b.lo[1])) * L1.px + 2 * (i - b.lo[0])];
}

// amrex::average_down of a cell array (velocity after the projection, gp: src/projection/
// incflo_apply_nodal_projection.cpp:258-266): covered coarse cell <- mean of its 8 fine children.
// Fine indices are relative to the fine box (cell 0 = first fine cell of the box).
__global__ void __launch_bounds__(256) k_comp_avgdown(const CBox b, Fab fine, Fab crse, int ncomp)
{
    const int i = b.lo[0] + blockIdx.x * 64 + (threadIdx.x & 63);
    const int j = b.lo[1] + blockIdx.y * 4 + (threadIdx.x >> 6);
    const int k = b.lo[2] + blockIdx.z;
    if (i > b.hi[0] || j > b.hi[1]) return;
    const int fi = 2 * (i - b.lo[0]), fj = 2 * (j - b.lo[1]), fk = 2 * (k - b.lo[2]);
    for (int c = 0; c < ncomp; ++c) {
        double s = 0.0;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc)
#pragma unroll
            for (int bb = 0; bb < 2; ++bb)
#pragma unroll
                for (int a = 0; a < 2; ++a) s += fine.p[fine.idx(fi + a, fj + bb, fk + cc, c)];
        crse.p[crse.idx(i, j, k, c)] = 0.125 * s;
    }
}

}  // namespace b200np_dev
