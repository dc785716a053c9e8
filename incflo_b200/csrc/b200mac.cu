// b200mac.cu -- the MAC projection behind the C ABI b200mac_* (include/b200np.h): Hydro::MacProjector over amrex::MLMG /
// MLABecLaplacian as incflo drives it in src/convection/incflo_compute_MAC_projected_velocities.cpp:69-129 and :280-299.
//
// Cell-centred, 7-point, variable face coefficient b = dt / rho:
//     A phi = - sum_d (1/dx_d^2) [ b_d(i+1/2) (phi(i+1) - phi(i)) - b_d(i-1/2) (phi(i) - phi(i-1)) ]        (mlabeclap_adotx)
// Boundary conditions (incflo::get_mac_projection_bc, src/projection/incflo_projection_bc.cpp:43-79) are evaluated on the
// fly, no ghost cells are stored: periodic wrap; Neumann ghost = first interior cell; Dirichlet (phi = 0 on the face)
// ghost = -2 phi_0 + phi_1 / 3, AMReX's maxorder = 3 extrapolation (mllinop_apply_bc).  Smoother: red-black Gauss-Seidel
// with over-relaxation 1.15 and the boundary stencil folded into the diagonal (abec_gsrb); restriction = mean of 8 cells,
// interpolation piecewise constant, coarse b = mean of the 4 coincident fine faces; MLMG V(2,2) cycle, BiCGStab bottom
// solve in one CTA.  Restated in oracle/mac_oracle.py (numpy), which tests/ compare this file with.
//
// HBM traffic per cell: a half-sweep reads phi (8 B), rhs of the active colour (4 B on average), three face arrays (24 B)
// and writes the active colour (4 B): 40 B/cell; residual 48 B/cell.  These kernels are plain one-thread-per-cell
// kernels with coalesced rows -- the 7-point stencil leaves its re-use to L1/L2 -- and are the "first correct path" of
// this operator; the nodal projection (b200np.cu) is the tuned path.
#include "../../include/b200np.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

enum { MAC_PER = 0, MAC_NEU = 1, MAC_DIR = 2 };
constexpr double MAC_OMEGA = 1.15;   // abec_gsrb

struct MacLev {
    int n[3];
    int bclo[3], bchi[3];
    double dh[3];          // 1 / dx^2
    double dxinv[3];
    double cflo[3], cfhi[3];   // coefficient of the first interior cell in the ghost formula: 0 periodic, 1 Neumann, -2 Dirichlet
    const double* b[3];    // face coefficients, dense: x (nz, ny, nx+1), y (nz, ny+1, nx), z (nz+1, ny, nx)
};

// caller array with its own box
struct MFab {
    double* p;
    int lo[3];
    int nx, ny;
    __host__ __device__ __forceinline__ long long idx(int i, int j, int k) const
    {
        return (i - lo[0]) + (long long)nx * ((j - lo[1]) + (long long)ny * (k - lo[2]));
    }
};

// Programmatic dependent launch: pdl_wait() blocks until the predecessor kernel has completed and its writes are visible (must precede
// the first access to data it produced); pdl_trigger() lets the successor become resident early (it waits in its own pdl_wait()).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

#define MCK(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            fprintf(stderr, "b200mac: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            throw int(B200NP_ERR_CUDA);                                                             \
        }                                                                                           \
    } while (0)

__device__ __forceinline__ long long cidx(const MacLev& L, int i, int j, int k) { return ((long long)k * L.n[1] + j) * L.n[0] + i; }
__device__ __forceinline__ double bface(const MacLev& L, int d, int i, int j, int k)   // face (i,j,k) of direction d (lower face of cell (i,j,k))
{
    if (d == 0) return L.b[0][((long long)k * L.n[1] + j) * (L.n[0] + 1) + i];
    if (d == 1) return L.b[1][((long long)k * (L.n[1] + 1) + j) * L.n[0] + i];
    return L.b[2][((long long)k * L.n[1] + j) * L.n[0] + i];
}
// phi of the neighbour of cell (i,j,k) one step s = -1 / +1 along d, BC ghost cells included
__device__ __forceinline__ double nb(const MacLev& L, const double* phi, int i, int j, int k, int d, int s, double pc)
{
    int q[3] = {i, j, k};
    const int c = q[d] + s, n = L.n[d];
    if (c >= 0 && c < n) { q[d] = c; return phi[cidx(L, q[0], q[1], q[2])]; }
    const int bc = s < 0 ? L.bclo[d] : L.bchi[d];
    if (bc == MAC_PER) { q[d] = s < 0 ? n - 1 : 0; return phi[cidx(L, q[0], q[1], q[2])]; }
    if (bc == MAC_NEU) return pc;
    q[d] -= s;   // second cell from the face
    return -2.0 * pc + phi[cidx(L, q[0], q[1], q[2])] * (1.0 / 3.0);
}

template <bool MAXR>
__device__ __forceinline__ double block_reduce(double v, double* sh)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double w = __shfl_xor_sync(0xffffffffu, v, o);
        v = MAXR ? fmax(v, w) : v + w;
    }
    const int tid_ = threadIdx.y * blockDim.x + threadIdx.x;
    const int lane = tid_ & 31, wid = tid_ >> 5, nw = (blockDim.x * blockDim.y + 31) >> 5;
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    if (wid == 0) {
        v = lane < nw ? sh[lane] : (MAXR ? 0.0 : 0.0);
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

// A phi at one cell and the pieces the smoother needs
__device__ __forceinline__ double adotx_cell(const MacLev& L, const double* __restrict__ phi, int i, int j, int k, double pc)
{
    if (i > 0 && i < L.n[0] - 1 && j > 0 && j < L.n[1] - 1 && k > 0 && k < L.n[2] - 1) {   // interior: no boundary logic, all loads up front
        const long long c = cidx(L, i, j, k), sy = L.n[0], sz = (long long)L.n[0] * L.n[1];
        const double xl = phi[c - 1], xh = phi[c + 1], yl = phi[c - sy], yh = phi[c + sy], zl = phi[c - sz], zh = phi[c + sz];
        const long long cx = ((long long)k * L.n[1] + j) * (L.n[0] + 1) + i, cy = ((long long)k * (L.n[1] + 1) + j) * L.n[0] + i;
        const double bxl = L.b[0][cx], bxh = L.b[0][cx + 1], byl = L.b[1][cy], byh = L.b[1][cy + sy], bzl = L.b[2][c], bzh = L.b[2][c + sz];
        return -(L.dh[0] * (bxh * (xh - pc) - bxl * (pc - xl)) + L.dh[1] * (byh * (yh - pc) - byl * (pc - yl)) + L.dh[2] * (bzh * (zh - pc) - bzl * (pc - zl)));
    }
    double y = 0.0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double bl = bface(L, d, i, j, k), bh = bface(L, d, i + (d == 0), j + (d == 1), k + (d == 2));
        const double lo = nb(L, phi, i, j, k, d, -1, pc), hi = nb(L, phi, i, j, k, d, +1, pc);
        y -= L.dh[d] * (bh * (hi - pc) - bl * (pc - lo));
    }
    return y;
}

// out = rhs - A phi; optional inf-norm partials (one entry per block).  Launch: block (64, 4), grid (ceil(nx / 64), ceil(ny / 4), nz).
__global__ void __launch_bounds__(256) k_mac_residual(const MacLev L, const double* __restrict__ phi, const double* __restrict__ rhs,
                                                      double* __restrict__ out, double* __restrict__ norm_partial)
{
    __shared__ double sh[34];
    const int i = blockIdx.x * 64 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = blockIdx.z;
    double r = 0.0;
    pdl_wait();
    if (i < L.n[0] && j < L.n[1]) {
        const long long t = cidx(L, i, j, k);
        const double pc = phi[t];
        r = rhs[t] - adotx_cell(L, phi, i, j, k, pc);
        if (out) out[t] = r;
    }
    if (norm_partial) {
        const double amax = block_reduce<true>(fabs(r), sh);
        if (threadIdx.x == 0 && threadIdx.y == 0) norm_partial[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = amax;
    }
}

// one red-black half-sweep in place (abec_gsrb).  old: the values the neighbours are read from -- phi itself (the six
// neighbours of a cell have the other colour), or a snapshot on a level with an odd periodic extent, where the wrap
// joins two cells of the same colour (AMReX reads those from ghost cells filled before the half-sweep).
__global__ void __launch_bounds__(256) k_mac_gsrb(const MacLev L, double* phi, const double* old /* may alias phi */,
                                                  const double* __restrict__ rhs, int redblack)
{
    // block (64, 4) over (i / 2, j), blockIdx.z = k
    const int ih = blockIdx.x * 64 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = blockIdx.z;
    const int i = 2 * ih + ((j + k + redblack) & 1);
    if (gridDim.x * gridDim.y * gridDim.z <= 1184u) pdl_trigger();   // a single wave: let the next half-sweep's blocks in right away
    pdl_wait();
    if (i >= L.n[0] || j >= L.n[1]) return;
    const long long c = cidx(L, i, j, k);
    const double pc = old[c];
    double gamma = 0.0, delta = 0.0, rho = 0.0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int q = d == 0 ? i : d == 1 ? j : k;
        const double bl = bface(L, d, i, j, k), bh = bface(L, d, i + (d == 0), j + (d == 1), k + (d == 2));
        const double lo = nb(L, old, i, j, k, d, -1, pc), hi = nb(L, old, i, j, k, d, +1, pc);
        gamma += L.dh[d] * (bl + bh);
        rho += L.dh[d] * (bl * lo + bh * hi);
        if (q == 0) delta += L.dh[d] * bl * L.cflo[d];
        if (q == L.n[d] - 1) delta += L.dh[d] * bh * L.cfhi[d];
    }
    const double res = rhs[c] - (gamma * pc - rho);
    phi[c] = pc + MAC_OMEGA / (gamma - delta) * res;
}

// crse = mean of the 8 fine cells (MLCellLinOp::restriction)
__global__ void __launch_bounds__(256) k_mac_restrict(const MacLev C, int fnx, int fny, const double* __restrict__ fine, double* __restrict__ crse)
{
    pdl_wait();
    const long long N = (long long)C.n[0] * C.n[1] * C.n[2];
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < N; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % C.n[0]), j = (int)((t / C.n[0]) % C.n[1]), k = (int)(t / ((long long)C.n[0] * C.n[1]));
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const double2 v = *reinterpret_cast<const double2*>(fine + ((long long)(2 * k + c) * fny + (2 * j + b)) * fnx + 2 * i);
                s += v.x + v.y;
            }
        crse[t] = 0.125 * s;
    }
}
// fine += crse(i/2, j/2, k/2) (MLCellLinOp::interpolation)
__global__ void __launch_bounds__(256) k_mac_interp_add(const MacLev F, int cnx, int cny, double* __restrict__ fine, const double* __restrict__ crse)
{
    pdl_wait();
    const long long N = (long long)F.n[0] * F.n[1] * F.n[2];
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < N; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % F.n[0]), j = (int)((t / F.n[0]) % F.n[1]), k = (int)(t / ((long long)F.n[0] * F.n[1]));
        fine[t] += crse[((long long)(k >> 1) * cny + (j >> 1)) * cnx + (i >> 1)];
    }
}
// coarse face coefficient = mean of the 4 coincident fine faces (amrex::average_down_faces), direction d
__global__ void __launch_bounds__(256) k_mac_coarsen_b(int d, int cn0, int cn1, int cn2, const double* __restrict__ fb, double* __restrict__ cb)
{
    const int ex = cn0 + (d == 0), ey = cn1 + (d == 1), ez = cn2 + (d == 2);          // coarse face array extents
    const int fx = 2 * cn0 + (d == 0), fy = 2 * cn1 + (d == 1);                       // fine face array extents (x, y)
    const long long N = (long long)ex * ey * ez;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < N; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % ex), j = (int)((t / ex) % ey), k = (int)(t / ((long long)ex * ey));
        double s = 0.0;
        for (int q = 0; q < 2; ++q)
            for (int p = 0; p < 2; ++p) {
                const int fi = 2 * i + (d == 0 ? 0 : p), fj = 2 * j + (d == 1 ? 0 : (d == 0 ? p : q)), fk = 2 * k + (d == 2 ? 0 : q);
                s += fb[((long long)fk * fy + fj) * fx + fi];
            }
        cb[t] = 0.25 * s;
    }
}
__global__ void __launch_bounds__(256) k_mac_fill(double* __restrict__ p, long long n, double v)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) p[t] = v;
}
// the caller's face coefficient array (own box) -> dense level-0 array
__global__ void __launch_bounds__(256) k_mac_copy_b(int ex, int ey, int ez, MFab src, double* __restrict__ dst)
{
    const long long N = (long long)ex * ey * ez;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < N; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % ex), j = (int)((t / ex) % ey), k = (int)(t / ((long long)ex * ey));
        dst[t] = src.p[src.idx(i, j, k)];
    }
}
// rhs = -div(u_mac) (MacProjector::project: computeDivergence, then mult(-1))
__global__ void __launch_bounds__(256) k_mac_div(const MacLev L, MFab u, MFab v, MFab w, double* __restrict__ rhs)
{
    const long long N = (long long)L.n[0] * L.n[1] * L.n[2];
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < N; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % L.n[0]), j = (int)((t / L.n[0]) % L.n[1]), k = (int)(t / ((long long)L.n[0] * L.n[1]));
        rhs[t] = -((u.p[u.idx(i + 1, j, k)] - u.p[u.idx(i, j, k)]) * L.dxinv[0] + (v.p[v.idx(i, j + 1, k)] - v.p[v.idx(i, j, k)]) * L.dxinv[1] +
                   (w.p[w.idx(i, j, k + 1)] - w.p[w.idx(i, j, k)]) * L.dxinv[2]);
    }
}
// u_mac += flux, flux = -b grad phi on every face of direction d, boundary faces with the BC ghost cell
// (MLMG::getFluxes(Location::FaceCenter) + MultiFab::Add in MacProjector::project)
__global__ void __launch_bounds__(256) k_mac_update(const MacLev L, int d, const double* __restrict__ phi, MFab u)
{
    const int ex = L.n[0] + (d == 0), ey = L.n[1] + (d == 1), ez = L.n[2] + (d == 2);
    const long long N = (long long)ex * ey * ez;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < N; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % ex), j = (int)((t / ex) % ey), k = (int)(t / ((long long)ex * ey));
        int q[3] = {i, j, k};
        double g;
        if (q[d] == L.n[d]) {       // high boundary face: the cell below it and its upper ghost
            q[d] -= 1;
            const double pc = phi[cidx(L, q[0], q[1], q[2])];
            g = nb(L, phi, q[0], q[1], q[2], d, +1, pc) - pc;
        } else {
            const double pc = phi[cidx(L, q[0], q[1], q[2])];
            g = pc - nb(L, phi, q[0], q[1], q[2], d, -1, pc);
        }
        u.p[u.idx(i, j, k)] -= bface(L, d, i, j, k) * g * L.dxinv[d];
    }
}
// phi copy-in / copy-out between the caller's cell array and the dense level-0 array
__global__ void __launch_bounds__(256) k_mac_copy_phi(const MacLev L, MFab f, double* __restrict__ dense, int to_dense)
{
    const long long N = (long long)L.n[0] * L.n[1] * L.n[2];
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < N; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % L.n[0]), j = (int)((t / L.n[0]) % L.n[1]), k = (int)(t / ((long long)L.n[0] * L.n[1]));
        if (to_dense) dense[t] = f.p[f.idx(i, j, k)]; else f.p[f.idx(i, j, k)] = dense[t];
    }
}
__global__ void __launch_bounds__(1024) k_mac_max_final(const double* __restrict__ partial, int nb_, double* __restrict__ out)
{
    __shared__ double sh[34];
    double a = 0.0;
    for (int t = threadIdx.x; t < nb_; t += blockDim.x) a = fmax(a, partial[t]);
    a = block_reduce<true>(a, sh);
    if (threadIdx.x == 0) out[0] = a;
}
__global__ void __launch_bounds__(256) k_mac_absmax_partial(const double* __restrict__ x, long long n, double* __restrict__ partial)
{
    __shared__ double sh[34];
    double a = 0.0;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) a = fmax(a, fabs(x[t]));
    a = block_reduce<true>(a, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = a;
}
// x -= mean(x): partial sums per block, then one CTA adds them -- a fixed order for a given size (solvability offset of all-Neumann /
// periodic problems)
__global__ void __launch_bounds__(256) k_mac_sum_partial(const double* __restrict__ x, long long n, double* __restrict__ partial)
{
    __shared__ double sh[34];
    double a = 0.0;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) a += x[t];
    a = block_reduce<false>(a, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = a;
}
__global__ void __launch_bounds__(1024) k_mac_sum(const double* __restrict__ partial, int nb_, long long n, double* __restrict__ out)
{
    __shared__ double sh[34];
    double a = 0.0;
    for (int t = threadIdx.x; t < nb_; t += blockDim.x) a += partial[t];
    a = block_reduce<false>(a, sh);
    if (threadIdx.x == 0) out[0] = a / (double)n;
}
__global__ void __launch_bounds__(256) k_mac_sub(double* __restrict__ x, long long n, const double* __restrict__ mean)
{
    const double m = mean[0];
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) x[t] -= m;
}
__global__ void __launch_bounds__(256) k_mac_axpy(double* __restrict__ y, const double* __restrict__ x, long long n)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) y[t] += x[t];
}

// Bottom solve in ONE CTA: MLMG::bottomSolve with MLCGSolver::solve_bicgstab (homogeneous BCs, plain dot products, reductions in
// a fixed order), solvability offset first when the operator is singular; on failure start over with 8 smooth calls.
// work: 8 vectors of N doubles.  info[0] += iterations, info[1] = return code.
__global__ void __launch_bounds__(1024) k_mac_bottom(const MacLev L, double* __restrict__ x, double* __restrict__ b, double* __restrict__ work,
                                                     int maxiter, double eps_rel, double eps_abs, int singular, int* __restrict__ info)
{
    __shared__ double sh[34];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int N = L.n[0] * L.n[1] * L.n[2];
    double *r = work, *rh = work + N, *p = work + 2 * (long long)N, *v = work + 3 * (long long)N, *s = work + 4 * (long long)N, *tt = work + 5 * (long long)N,
           *snap = work + 6 * (long long)N;
    auto cell = [&](int t, int& i, int& j, int& k) { i = t % L.n[0]; j = (t / L.n[0]) % L.n[1]; k = t / (L.n[0] * L.n[1]); };
    auto apply = [&](const double* in, double* out) {   // out = A in
        __syncthreads();
        for (int t = tid; t < N; t += nt) { int i, j, k; cell(t, i, j, k); out[t] = adotx_cell(L, in, i, j, k, in[t]); }
        __syncthreads();
    };
    auto dot = [&](const double* a, const double* c) { double q = 0.0; for (int t = tid; t < N; t += nt) q += a[t] * c[t]; return block_reduce<false>(q, sh); };
    auto ninf = [&](const double* a) { double q = 0.0; for (int t = tid; t < N; t += nt) q = fmax(q, fabs(a[t])); return block_reduce<true>(q, sh); };
    if (singular) {
        double q = 0.0;
        for (int t = tid; t < N; t += nt) q += b[t];
        const double mean = block_reduce<false>(q, sh) / (double)N;
        for (int t = tid; t < N; t += nt) b[t] -= mean;
        __syncthreads();
    }
    // x == 0 on entry: r = b
    for (int t = tid; t < N; t += nt) { r[t] = b[t]; rh[t] = b[t]; x[t] = 0.0; }
    __syncthreads();
    double rnorm = ninf(r);
    const double rnorm0 = rnorm;
    int ret = 0, it = 0;
    if (!(rnorm0 == 0.0 || rnorm0 < eps_abs)) {
        double rho_1 = 0.0, alpha = 0.0, omega = 0.0;
        bool done = false;
        for (it = 1; it <= maxiter && !done; ++it) {
            const double rho = dot(rh, r);
            if (rho == 0.0) { ret = 1; break; }
            if (it == 1) { for (int t = tid; t < N; t += nt) p[t] = r[t]; }
            else {
                const double beta = (rho / rho_1) * (alpha / omega);
                for (int t = tid; t < N; t += nt) p[t] = r[t] + beta * (p[t] - omega * v[t]);
            }
            apply(p, v);
            const double rhTv = dot(rh, v);
            if (rhTv == 0.0) { ret = 2; break; }
            alpha = rho / rhTv;
            for (int t = tid; t < N; t += nt) { x[t] += alpha * p[t]; s[t] = r[t] - alpha * v[t]; }
            __syncthreads();
            rnorm = ninf(s);
            if (rnorm < eps_rel * rnorm0 || rnorm < eps_abs) { done = true; break; }
            apply(s, tt);
            const double t2 = dot(tt, tt);
            if (t2 == 0.0) { ret = 3; break; }
            omega = dot(tt, s) / t2;
            for (int t = tid; t < N; t += nt) { x[t] += omega * s[t]; r[t] = s[t] - omega * tt[t]; }
            __syncthreads();
            rnorm = ninf(r);
            if (rnorm < eps_rel * rnorm0 || rnorm < eps_abs) { done = true; break; }
            if (omega == 0.0) { ret = 4; break; }
            rho_1 = rho;
        }
        if (ret == 0 && !(rnorm < eps_rel * rnorm0 || rnorm < eps_abs)) ret = 8;
        if (it > maxiter) it = maxiter;
    }
    __syncthreads();
    if (ret != 0 && ret != 8) { for (int t = tid; t < N; t += nt) x[t] = 0.0; }
    if (ret != 0) {   // MLMG::bottomSolve: start over with nuf = 8 smooth calls
        for (int t = tid; t < N; t += nt) x[t] = 0.0;
        __syncthreads();
        for (int call = 0; call < 8; ++call)
            for (int rb = 0; rb < 2; ++rb) {
                for (int t = tid; t < N; t += nt) snap[t] = x[t];
                __syncthreads();
                for (int t = tid; t < N; t += nt) {
                    int i, j, k; cell(t, i, j, k);
                    if (((i + j + k + rb) & 1) != 0) continue;
                    const double pc = snap[t];
                    double gamma = 0.0, delta = 0.0, rho = 0.0;
                    for (int d = 0; d < 3; ++d) {
                        const int q = d == 0 ? i : d == 1 ? j : k;
                        const double bl = bface(L, d, i, j, k), bh = bface(L, d, i + (d == 0), j + (d == 1), k + (d == 2));
                        gamma += L.dh[d] * (bl + bh);
                        rho += L.dh[d] * (bl * nb(L, snap, i, j, k, d, -1, pc) + bh * nb(L, snap, i, j, k, d, +1, pc));
                        if (q == 0) delta += L.dh[d] * bl * L.cflo[d];
                        if (q == L.n[d] - 1) delta += L.dh[d] * bh * L.cfhi[d];
                    }
                    x[t] = pc + MAC_OMEGA / (gamma - delta) * (b[t] - (gamma * pc - rho));
                }
                __syncthreads();
            }
    }
    if (tid == 0) { atomicAdd(info, it); info[1] = ret; }
}

// ---- multi-box MultiFabs (every reference deck: amr.max_grid_size = 16): gather into / scatter from one dense array per field ----
struct MacMfFab {
    double* p;
    int lo[3];           // allocated box
    int nx, ny, nz;
};
// dense (ex, ey, ez) <- the valid boxes of the fabs (allocated box shrunk by ngrow; for a face-centred MultiFab that is the cells' box
// plus the far face, which adjacent boxes share: both hold the same value)
__global__ void __launch_bounds__(256) k_mac_mf_gather(const MacMfFab* __restrict__ tab, int ngrow, double* __restrict__ dense, int ex, int ey, int ez)
{
    const MacMfFab f = tab[blockIdx.y];
    const int vx = f.nx - 2 * ngrow, vy = f.ny - 2 * ngrow, vz = f.nz - 2 * ngrow;
    const long long total = (long long)vx * vy * vz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int li = (int)(t % vx) + ngrow, lj = (int)((t / vx) % vy) + ngrow, lk = (int)(t / ((long long)vx * vy)) + ngrow;
        const int i = li + f.lo[0], j = lj + f.lo[1], k = lk + f.lo[2];
        if (i < 0 || i >= ex || j < 0 || j >= ey || k < 0 || k >= ez) continue;
        dense[((long long)k * ey + j) * ex + i] = f.p[((long long)lk * f.ny + lj) * f.nx + li];
    }
}
__global__ void __launch_bounds__(256) k_mac_mf_scatter(const MacMfFab* __restrict__ tab, int ngrow, const double* __restrict__ dense, int ex, int ey, int ez)
{
    const MacMfFab f = tab[blockIdx.y];
    const int vx = f.nx - 2 * ngrow, vy = f.ny - 2 * ngrow, vz = f.nz - 2 * ngrow;
    const long long total = (long long)vx * vy * vz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int li = (int)(t % vx) + ngrow, lj = (int)((t / vx) % vy) + ngrow, lk = (int)(t / ((long long)vx * vy)) + ngrow;
        const int i = li + f.lo[0], j = lj + f.lo[1], k = lk + f.lo[2];
        if (i < 0 || i >= ex || j < 0 || j >= ey || k < 0 || k >= ez) continue;
        f.p[((long long)lk * f.ny + lj) * f.nx + li] = dense[((long long)k * ey + j) * ex + i];
    }
}

struct MacLevel {
    MacLev g{};
    double* b[3] = {nullptr, nullptr, nullptr};
    double *cor = nullptr, *res = nullptr, *rescor = nullptr, *sol = nullptr, *rhs = nullptr;
    long long ncell = 0;
    bool odd_periodic = false;   // the wrap joins two cells of one colour: the half-sweeps read a snapshot
};

bool is_dev_ptr(const void* p)
{
    if (!p) return true;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
size_t box_doubles(const b200np_fab* b) { return (size_t)(b->hi[0] - b->lo[0] + 1) * (b->hi[1] - b->lo[1] + 1) * (b->hi[2] - b->lo[2] + 1); }
MFab mfab(double* p, const b200np_fab* b)
{
    MFab f{};
    f.p = p;
    if (!b) return f;
    for (int d = 0; d < 3; ++d) f.lo[d] = b->lo[d];
    f.nx = b->hi[0] - b->lo[0] + 1; f.ny = b->hi[1] - b->lo[1] + 1;
    return f;
}
dim3 mac_grid3(const MacLev& g, int xdiv = 1) { return dim3(((g.n[0] + xdiv - 1) / xdiv + 63) / 64, (g.n[1] + 3) / 4, g.n[2]); }
int mac_blocks3(const MacLev& g) { const dim3 d = mac_grid3(g); return (int)(d.x * d.y * d.z); }
int grid_for(long long n) { return (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, 148 * 8)); }

}  // namespace

struct b200mac {
    b200np_geom geom{};
    b200np_opts opts{};
    int device = 0;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    std::vector<MacLevel> lv;
    std::vector<void*> allocs;
    double *partial = nullptr, *dscal = nullptr, *work = nullptr, *snap = nullptr;
    int* dinfo = nullptr;
    double* hscal = nullptr;
    int* hinfo = nullptr;
    bool singular = true, have_coeffs = false;
    bool top_direct = true;   // finest level of the V-cycle relaxes (sol, rhs) directly (B200MAC_TOP_DIRECT=0: MLMG's correction form)
    long long launches = 0;
    struct Stage { double* d = nullptr; size_t bytes = 0; } stage[8];
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    struct Mf { double* dense = nullptr; size_t dense_doubles = 0; double* stage = nullptr; size_t stage_doubles = 0; MacMfFab* tab = nullptr; size_t tab_cap = 0; } mf[8];
    cudaGraph_t graph = nullptr;          // one V-cycle, captured once (every pointer in it belongs to the handle)
    cudaGraphExec_t graph_exec = nullptr;
    long long launches_per_vcycle = 0;
};

namespace {

#define MLAUNCH(h, kern, grid, block, ...)                 \
    do {                                                   \
        kern<<<grid, block, 0, (h)->stream>>>(__VA_ARGS__); \
        (h)->launches++;                                   \
    } while (0)

// launch with the programmatic-stream-serialization attribute: ONLY for kernels that call pdl_wait() before touching data
template <typename... KArgs, typename... Args>
void mac_launch_pdl(b200mac* h, void (*kern)(KArgs...), dim3 grid, dim3 block, Args... args)
{
    static const int use_pdl = getenv("B200MAC_PDL") ? atoi(getenv("B200MAC_PDL")) : 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = use_pdl ? 1 : 0;
    MCK(cudaLaunchKernelEx(&cfg, kern, KArgs(args)...));
    h->launches++;
}

double* mac_alloc(b200mac* h, size_t doubles)
{
    void* p = nullptr;
    MCK(cudaMalloc(&p, std::max<size_t>(doubles, 1) * sizeof(double)));
    MCK(cudaMemsetAsync(p, 0, std::max<size_t>(doubles, 1) * sizeof(double), h->stream));
    h->allocs.push_back(p);
    return static_cast<double*>(p);
}

void mac_build(b200mac* h)
{
    const b200np_geom& G = h->geom;
    int n[3] = {G.n_cell[0], G.n_cell[1], G.n_cell[2]};
    double dx[3] = {G.dx[0], G.dx[1], G.dx[2]};
    h->singular = true;
    for (int lev = 0;; ++lev) {
        MacLevel L;
        for (int d = 0; d < 3; ++d) {
            auto bc = [](int b) { return b == B200NP_BC_PERIODIC ? MAC_PER : b == B200NP_BC_DIRICHLET ? MAC_DIR : MAC_NEU; };   // inflow -> Neumann (:60-66)
            L.g.n[d] = n[d]; L.g.bclo[d] = bc(G.bc_lo[d]); L.g.bchi[d] = bc(G.bc_hi[d]);
            L.g.dh[d] = 1.0 / (dx[d] * dx[d]); L.g.dxinv[d] = 1.0 / dx[d];
            auto cf = [&](int b) { return b == MAC_PER ? 0.0 : b == MAC_NEU ? 1.0 : (std::min(n[d] + 1, 3) >= 3 ? -2.0 : -1.0); };
            L.g.cflo[d] = cf(L.g.bclo[d]); L.g.cfhi[d] = cf(L.g.bchi[d]);
            if (L.g.bclo[d] == MAC_DIR || L.g.bchi[d] == MAC_DIR) h->singular = false;
            if (L.g.bclo[d] == MAC_PER && (n[d] & 1)) L.odd_periodic = true;
        }
        L.ncell = (long long)n[0] * n[1] * n[2];
        L.b[0] = mac_alloc(h, (size_t)(n[0] + 1) * n[1] * n[2]);
        L.b[1] = mac_alloc(h, (size_t)n[0] * (n[1] + 1) * n[2]);
        L.b[2] = mac_alloc(h, (size_t)n[0] * n[1] * (n[2] + 1));
        for (int d = 0; d < 3; ++d) L.g.b[d] = L.b[d];
        L.cor = mac_alloc(h, L.ncell); L.res = mac_alloc(h, L.ncell); L.rescor = mac_alloc(h, L.ncell);
        if (lev == 0) { L.sol = mac_alloc(h, L.ncell); L.rhs = mac_alloc(h, L.ncell); }
        h->lv.push_back(L);
        bool can = lev + 1 <= h->opts.mg_max_coarsening_level && lev + 1 < 30;
        for (int d = 0; d < 3; ++d) if (n[d] % 2 != 0 || n[d] / 2 < 2) can = false;
        if (!can) break;
        for (int d = 0; d < 3; ++d) { n[d] /= 2; dx[d] *= 2; }
    }
    const MacLevel& B = h->lv.back();
    h->work = mac_alloc(h, (size_t)8 * B.ncell);
    h->snap = mac_alloc(h, (size_t)h->lv[0].ncell);
    h->partial = mac_alloc(h, std::max(148 * 8, mac_blocks3(h->lv[0].g)) + 8);
    h->dscal = mac_alloc(h, 16);
    h->dinfo = reinterpret_cast<int*>(mac_alloc(h, 4));
    MCK(cudaMallocHost(&h->hscal, 16 * sizeof(double)));
    MCK(cudaMallocHost(&h->hinfo, 8 * sizeof(int)));
}

void mac_coarsen_coeffs(b200mac* h)
{
    for (size_t l = 0; l + 1 < h->lv.size(); ++l) {
        MacLevel &F = h->lv[l], &C = h->lv[l + 1];
        for (int d = 0; d < 3; ++d) {
            const long long nf = (long long)(C.g.n[0] + (d == 0)) * (C.g.n[1] + (d == 1)) * (C.g.n[2] + (d == 2));
            MLAUNCH(h, k_mac_coarsen_b, grid_for(nf), 256, d, C.g.n[0], C.g.n[1], C.g.n[2], (const double*)F.b[d], C.b[d]);
        }
    }
}

// one MLMG smooth call: red half-sweep, black half-sweep
void mac_smooth(b200mac* h, MacLevel& L, double* phi, const double* rhs, int ncalls)
{
    for (int c = 0; c < ncalls; ++c)
        for (int rb = 0; rb < 2; ++rb) {
            const double* old = phi;
            if (L.odd_periodic) {
                MCK(cudaMemcpyAsync(h->snap, phi, L.ncell * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
                old = h->snap;
            }
            mac_launch_pdl(h, k_mac_gsrb, mac_grid3(L.g, 2), dim3(64, 4), L.g, phi, old, rhs, rb);
        }
}

// One V-cycle.  top_direct (default): the finest level relaxes (sol, rhs) in place instead of (cor, res = rhs - A sol) from cor = 0 -- the
// smoother is a stationary linear iteration, so both forms do the same arithmetic up to rounding, and the direct one needs no zeroed cor,
// no sol += cor pass (24 B/cell) and no stored top-level residual (8 B/cell): the caller only takes the norm (as b200np.cu's V-cycle does).
void mac_vcycle(b200mac* h)
{
    const int nl = (int)h->lv.size();
    const int nu1 = h->opts.num_pre_smooth, nu2 = h->opts.num_post_smooth;
    for (int l = 0; l < nl - 1; ++l) {
        MacLevel &L = h->lv[l], &C = h->lv[l + 1];
        const bool direct = l == 0 && h->top_direct;   // (nl > 1 here)
        double* x = direct ? L.sol : L.cor;
        const double* b = direct ? L.rhs : L.res;
        if (!direct) MCK(cudaMemsetAsync(L.cor, 0, L.ncell * sizeof(double), h->stream));
        mac_smooth(h, L, x, b, nu1);
        mac_launch_pdl(h, k_mac_residual, mac_grid3(L.g), dim3(64, 4), L.g, (const double*)x, b, L.rescor, (double*)nullptr);
        mac_launch_pdl(h, k_mac_restrict, dim3(grid_for(C.ncell)), dim3(256), C.g, L.g.n[0], L.g.n[1], (const double*)L.rescor, C.res);
    }
    MacLevel& B = h->lv.back();
    MLAUNCH(h, k_mac_bottom, 1, 1024, B.g, B.cor, B.res, h->work, h->opts.bottom_maxiter, h->opts.bottom_rtol, h->opts.bottom_atol,
            h->singular ? 1 : 0, h->dinfo);
    for (int l = nl - 2; l >= 0; --l) {
        MacLevel &L = h->lv[l], &C = h->lv[l + 1];
        const bool direct = l == 0 && h->top_direct;
        double* x = direct ? L.sol : L.cor;
        const double* b = direct ? L.rhs : L.res;
        mac_launch_pdl(h, k_mac_interp_add, dim3(grid_for(L.ncell)), dim3(256), L.g, C.g.n[0], C.g.n[1], x, (const double*)C.cor);
        mac_smooth(h, L, x, b, nu2);
    }
}

// the V-cycle as a CUDA graph: ~80 small launches per cycle, most of them on levels that are pure launch latency
void mac_vcycle_run(b200mac* h)
{
    if (!h->opts.use_graph) { mac_vcycle(h); return; }
    if (!h->graph_exec) {
        const long long before = h->launches;
        MCK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        try { mac_vcycle(h); }
        catch (int) {
            cudaGraph_t broken = nullptr;
            cudaStreamEndCapture(h->stream, &broken);
            if (broken) cudaGraphDestroy(broken);
            cudaGetLastError();
            throw;
        }
        MCK(cudaStreamEndCapture(h->stream, &h->graph));
        MCK(cudaGraphInstantiate(&h->graph_exec, h->graph, 0));
        h->launches_per_vcycle = h->launches - before;
        h->launches = before;
    }
    MCK(cudaGraphLaunch(h->graph_exec, h->stream));
    h->launches += h->launches_per_vcycle;
}

double mac_read_norm(b200mac* h, int nb_)
{
    MLAUNCH(h, k_mac_max_final, 1, 1024, (const double*)h->partial, nb_, h->dscal + 2);
    MCK(cudaMemcpyAsync(h->hscal + 2, h->dscal + 2, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    MCK(cudaStreamSynchronize(h->stream));
    return h->hscal[2];
}

// MLMG::solve on (sol, rhs) of level 0
int mac_solve(b200mac* h, double rtol, double atol, b200np_stats* st)
{
    MacLevel& L0 = h->lv[0];
    st->iters = 0; st->bottom_iters = 0; st->status = B200NP_OK; st->nlevels = (int)h->lv.size();
    MCK(cudaMemsetAsync(h->dinfo, 0, 4 * sizeof(int), h->stream));
    if (h->singular) {   // makeSolvable: remove the mean of rhs
        MLAUNCH(h, k_mac_sum_partial, grid_for(L0.ncell), 256, (const double*)L0.rhs, L0.ncell, h->partial);
        MLAUNCH(h, k_mac_sum, 1, 1024, (const double*)h->partial, grid_for(L0.ncell), L0.ncell, h->dscal);
        MLAUNCH(h, k_mac_sub, grid_for(L0.ncell), 256, L0.rhs, L0.ncell, (const double*)h->dscal);
    }
    const int nb_ = grid_for(L0.ncell);
    MLAUNCH(h, k_mac_absmax_partial, nb_, 256, (const double*)L0.rhs, L0.ncell, h->partial);
    st->rhsnorm = mac_read_norm(h, nb_);
    const bool direct = h->top_direct && h->lv.size() > 1;   // a single level: the bottom solve is the cycle, on (cor, res)
    double* const top_res = direct ? (double*)nullptr : L0.res;   // direct form: only the norm of the top-level residual is needed
    MLAUNCH(h, k_mac_residual, mac_grid3(L0.g), dim3(64, 4), L0.g, (const double*)L0.sol, (const double*)L0.rhs, top_res, h->partial);
    st->resnorm0 = mac_read_norm(h, mac_blocks3(L0.g));
    const double maxnorm = std::max(st->rhsnorm, st->resnorm0);
    const double target = std::max(atol, std::max(rtol, 1e-16) * maxnorm);
    st->resnorm = st->resnorm0;
    st->resnorm_hist[0] = st->resnorm0;
    if (h->opts.verbose >= 1) printf("MLMG: Initial rhs               = %.12g\nMLMG: Initial residual (resid0) = %.12g\n", st->rhsnorm, st->resnorm0);
    if (st->resnorm0 <= target) return B200NP_OK;
    bool converged = false;
    for (int it = 0; it < h->opts.maxiter; ++it) {
        mac_vcycle_run(h);
        if (!direct) MLAUNCH(h, k_mac_axpy, nb_, 256, L0.sol, (const double*)L0.cor, L0.ncell);
        MLAUNCH(h, k_mac_residual, mac_grid3(L0.g), dim3(64, 4), L0.g, (const double*)L0.sol, (const double*)L0.rhs, top_res, h->partial);
        st->resnorm = mac_read_norm(h, mac_blocks3(L0.g));
        st->iters = it + 1;
        if (it + 1 < 128) st->resnorm_hist[it + 1] = st->resnorm;
        if (h->opts.verbose >= 2) printf("MLMG: Iteration %3d Fine resid/bnorm = %.12g\n", it + 1, st->resnorm / maxnorm);
        if (st->resnorm <= target) { converged = true; break; }
        if (!(st->resnorm <= 1e20 * maxnorm)) { st->status = B200NP_ERR_DIVERGED; break; }
    }
    if (!converged && st->status == B200NP_OK) st->status = B200NP_ERR_NOT_CONVERGED;
    MCK(cudaMemcpyAsync(h->hinfo, h->dinfo, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    MCK(cudaStreamSynchronize(h->stream));
    st->bottom_iters = h->hinfo[0];
    if (h->opts.verbose >= 1) printf("MLMG: Final Iter. %d resid, resid/bnorm = %.12g, %.12g\n", st->iters, st->resnorm, st->resnorm / maxnorm);
    return st->status;
}

double* mac_stage_in(b200mac* h, int slot, const double* p, const b200np_fab* box, bool copy, bool* staged, b200np_stats* st)
{
    *staged = false;
    if (!p) return nullptr;
    if (is_dev_ptr(p)) return const_cast<double*>(p);
    const size_t bytes = box_doubles(box) * sizeof(double);
    auto& S = h->stage[slot];
    if (S.bytes < bytes) {
        if (S.d) MCK(cudaFree(S.d));
        MCK(cudaMalloc(&S.d, bytes));
        S.bytes = bytes;
    }
    if (copy) { MCK(cudaMemcpyAsync(S.d, p, bytes, cudaMemcpyHostToDevice, h->stream)); st->h2d_bytes += (long long)bytes; }
    *staged = true;
    return S.d;
}
void mac_stage_out(b200mac* h, int slot, double* p, const b200np_fab* box, bool staged, b200np_stats* st)
{
    if (!staged || !p) return;
    const size_t bytes = box_doubles(box) * sizeof(double);
    MCK(cudaMemcpyAsync(p, h->stage[slot].d, bytes, cudaMemcpyDeviceToHost, h->stream));
    st->d2h_bytes += (long long)bytes;
}
// the box must hold the index range [0, hi] of the array kind
bool mac_box_ok(const b200np_fab* b, int hx, int hy, int hz)
{
    if (!b) return false;
    return b->lo[0] <= 0 && b->lo[1] <= 0 && b->lo[2] <= 0 && b->hi[0] >= hx && b->hi[1] >= hy && b->hi[2] >= hz;
}

}  // namespace

extern "C" {

int b200mac_create(b200mac_t** out, const b200np_geom* geom, const b200np_opts* opts, int device)
{
    if (!out || !geom) return B200NP_ERR_BAD_ARG;
    *out = nullptr;
    for (int d = 0; d < 3; ++d) {
        if (geom->n_cell[d] < 2 || !(geom->dx[d] > 0)) return B200NP_ERR_BAD_ARG;
        if (geom->bc_lo[d] < 0 || geom->bc_lo[d] > 3 || geom->bc_hi[d] < 0 || geom->bc_hi[d] > 3) return B200NP_ERR_BAD_BC;
        if ((geom->bc_lo[d] == B200NP_BC_PERIODIC) != (geom->bc_hi[d] == B200NP_BC_PERIODIC)) return B200NP_ERR_BAD_BC;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) { cudaGetLastError(); return B200NP_ERR_CUDA; }
    b200mac* h = new b200mac();
    try {
        MCK(cudaSetDevice(device));
        h->device = device;
        h->geom = *geom;
        if (opts) h->opts = *opts;
        else { b200np_default_opts(&h->opts); h->opts.maxiter = 200; h->opts.bottom_maxiter = 200; }   // MLMG defaults (mac_proj.* keys)
        MCK(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
        h->stream = h->own_stream;
        for (auto& e : h->ev) MCK(cudaEventCreate(&e));
        if (const char* e = getenv("B200MAC_TOP_DIRECT")) h->top_direct = atoi(e) != 0;
        mac_build(h);
        MCK(cudaStreamSynchronize(h->stream));
    } catch (int e) { b200mac_destroy(h); return e; }
    *out = h;
    return B200NP_OK;
}

void b200mac_destroy(b200mac_t* h)
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
int b200mac_set_stream(b200mac_t* h, void* stream)
{
    if (!h) return B200NP_ERR_BAD_ARG;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    h->stream = stream ? (cudaStream_t)stream : h->own_stream;
    return B200NP_OK;
}

int b200mac_nlevels(const b200mac_t* h) { return h ? (int)h->lv.size() : 0; }

int b200mac_set_coeffs(b200mac_t* h, const double* bx, const b200np_fab* bx_box, const double* by, const b200np_fab* by_box, const double* bz,
                       const b200np_fab* bz_box, double const_beta)
{
    if (!h) return B200NP_ERR_BAD_ARG;
    if ((bx != nullptr) != (by != nullptr) || (bx != nullptr) != (bz != nullptr)) return B200NP_ERR_BAD_ARG;
    try {
        MCK(cudaSetDevice(h->device));
        MacLevel& L0 = h->lv[0];
        const int* n = L0.g.n;
        if (bx) {
            if (!mac_box_ok(bx_box, n[0], n[1] - 1, n[2] - 1) || !mac_box_ok(by_box, n[0] - 1, n[1], n[2] - 1) || !mac_box_ok(bz_box, n[0] - 1, n[1] - 1, n[2]))
                return B200NP_ERR_BAD_ARG;
            const double* src[3] = {bx, by, bz};
            const b200np_fab* box[3] = {bx_box, by_box, bz_box};
            b200np_stats st{};
            for (int d = 0; d < 3; ++d) {
                bool staged;
                double* dp = mac_stage_in(h, d, src[d], box[d], true, &staged, &st);
                const int ex = n[0] + (d == 0), ey = n[1] + (d == 1), ez = n[2] + (d == 2);
                MLAUNCH(h, k_mac_copy_b, grid_for((long long)ex * ey * ez), 256, ex, ey, ez, mfab(dp, box[d]), L0.b[d]);
            }
        } else {
            if (!(const_beta > 0)) return B200NP_ERR_BAD_ARG;
            for (int d = 0; d < 3; ++d) {
                const long long nf = (long long)(n[0] + (d == 0)) * (n[1] + (d == 1)) * (n[2] + (d == 2));
                MLAUNCH(h, k_mac_fill, grid_for(nf), 256, L0.b[d], nf, const_beta);
            }
        }
        mac_coarsen_coeffs(h);
        MCK(cudaStreamSynchronize(h->stream));
        h->have_coeffs = true;
        return B200NP_OK;
    } catch (int e) { return e; }
}

int b200mac_project(b200mac_t* h, double* umac, const b200np_fab* u_box, double* vmac, const b200np_fab* v_box, double* wmac,
                    const b200np_fab* w_box, double* mac_phi, const b200np_fab* phi_box, int phi_is_initial_guess, double rtol, double atol,
                    b200np_stats* stats)
{
    b200np_stats local{};
    b200np_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof(*st));
    if (!h || !umac || !vmac || !wmac || !h->have_coeffs) return st->status = B200NP_ERR_BAD_ARG;
    try {
        MCK(cudaSetDevice(h->device));
        MacLevel& L0 = h->lv[0];
        const int* n = L0.g.n;
        if (!mac_box_ok(u_box, n[0], n[1] - 1, n[2] - 1) || !mac_box_ok(v_box, n[0] - 1, n[1], n[2] - 1) || !mac_box_ok(w_box, n[0] - 1, n[1] - 1, n[2]))
            return st->status = B200NP_ERR_BAD_ARG;
        if (mac_phi && !mac_box_ok(phi_box, n[0] - 1, n[1] - 1, n[2] - 1)) return st->status = B200NP_ERR_BAD_ARG;
        h->launches = 0;
        MCK(cudaEventRecord(h->ev[0], h->stream));
        bool su, sv, sw, sp = false;
        double* du = mac_stage_in(h, 3, umac, u_box, true, &su, st);
        double* dv = mac_stage_in(h, 4, vmac, v_box, true, &sv, st);
        double* dw = mac_stage_in(h, 5, wmac, w_box, true, &sw, st);
        double* dp = mac_phi ? mac_stage_in(h, 6, mac_phi, phi_box, phi_is_initial_guess != 0, &sp, st) : nullptr;
        const MFab fu = mfab(du, u_box), fv = mfab(dv, v_box), fw = mfab(dw, w_box), fp = mfab(dp, phi_box);
        const int nb_ = grid_for(L0.ncell);
        MLAUNCH(h, k_mac_div, nb_, 256, L0.g, fu, fv, fw, L0.rhs);
        if (dp && phi_is_initial_guess) MLAUNCH(h, k_mac_copy_phi, nb_, 256, L0.g, fp, L0.sol, 1);
        else MCK(cudaMemsetAsync(L0.sol, 0, L0.ncell * sizeof(double), h->stream));
        MCK(cudaEventRecord(h->ev[2], h->stream));
        const int status = mac_solve(h, rtol, atol, st);
        MCK(cudaEventRecord(h->ev[3], h->stream));
        const MFab* f[3] = {&fu, &fv, &fw};
        for (int d = 0; d < 3; ++d) {
            const long long nf = (long long)(n[0] + (d == 0)) * (n[1] + (d == 1)) * (n[2] + (d == 2));
            MLAUNCH(h, k_mac_update, grid_for(nf), 256, L0.g, d, (const double*)L0.sol, *f[d]);
        }
        if (dp) MLAUNCH(h, k_mac_copy_phi, nb_, 256, L0.g, fp, L0.sol, 0);
        mac_stage_out(h, 3, umac, u_box, su, st); mac_stage_out(h, 4, vmac, v_box, sv, st); mac_stage_out(h, 5, wmac, w_box, sw, st);
        mac_stage_out(h, 6, mac_phi, phi_box, sp, st);
        MCK(cudaEventRecord(h->ev[1], h->stream));
        MCK(cudaEventSynchronize(h->ev[1]));
        float ms = 0;
        MCK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1])); st->ms_total = ms;
        MCK(cudaEventElapsedTime(&ms, h->ev[2], h->ev[3])); st->ms_solve = ms;
        st->launches = h->launches;
        return st->status = status;
    } catch (int e) { return st->status = e; }
}

}  // extern "C"

namespace {
// one MultiFab of the call: validated, its fabs on the device (staged if host pointers), gathered into a dense array
struct MacMfView {
    const b200np_mfab* m = nullptr;
    int slot = 0, ext[3] = {0, 0, 0};
    bool staged = false;
    std::vector<size_t> off;
    double* dense = nullptr;
    b200np_fab dense_box{};
};
// the valid boxes (allocated box shrunk by ngrow) minus the far face in direction d must tile the domain's cells
bool mac_mf_ok(const b200np_mfab* m, const int n[3], int d)
{
    if (!m || m->nfabs < 1 || m->ngrow < 0 || !m->box || !m->data) return false;
    long long vol = 0;
    for (int f = 0; f < m->nfabs; ++f) {
        if (!m->data[f]) return false;
        long long v = 1;
        for (int q = 0; q < 3; ++q) {
            const int vlo = m->box[f].lo[q] + m->ngrow, vhi = m->box[f].hi[q] - m->ngrow - (q == d ? 1 : 0);
            if (vhi < vlo || vlo < 0 || vhi > n[q] - 1) return false;
            v *= vhi - vlo + 1;
        }
        vol += v;
    }
    return vol == (long long)n[0] * n[1] * n[2];
}
void mac_mf_map(b200mac* h, MacMfView& V, int slot, const b200np_mfab* m, const int n[3], int d, bool copy_in, b200np_stats* st)
{
    V.m = m; V.slot = slot;
    for (int q = 0; q < 3; ++q) V.ext[q] = n[q] + (q == d ? 1 : 0);
    auto& S = h->mf[slot];
    const int nf = m->nfabs;
    V.staged = !is_dev_ptr(m->data[0]);
    V.off.assign(nf, 0);
    size_t total = 0;
    std::vector<MacMfFab> host(nf);
    for (int f = 0; f < nf; ++f) {
        const b200np_fab& b = m->box[f];
        host[f].nx = b.hi[0] - b.lo[0] + 1; host[f].ny = b.hi[1] - b.lo[1] + 1; host[f].nz = b.hi[2] - b.lo[2] + 1;
        for (int q = 0; q < 3; ++q) host[f].lo[q] = b.lo[q];
        V.off[f] = total;
        total += (size_t)host[f].nx * host[f].ny * host[f].nz;
    }
    if (V.staged) {
        if (S.stage_doubles < total) { if (S.stage) MCK(cudaFree(S.stage)); MCK(cudaMalloc(&S.stage, total * sizeof(double))); S.stage_doubles = total; }
        if (copy_in) {
            for (int f = 0; f < nf; ++f)
                MCK(cudaMemcpyAsync(S.stage + V.off[f], m->data[f], (size_t)host[f].nx * host[f].ny * host[f].nz * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            st->h2d_bytes += (long long)(total * sizeof(double));
        }
    }
    for (int f = 0; f < nf; ++f) host[f].p = V.staged ? S.stage + V.off[f] : m->data[f];
    if (S.tab_cap < (size_t)nf) { if (S.tab) MCK(cudaFree(S.tab)); MCK(cudaMalloc(&S.tab, (size_t)nf * sizeof(MacMfFab))); S.tab_cap = nf; }
    MCK(cudaMemcpyAsync(S.tab, host.data(), (size_t)nf * sizeof(MacMfFab), cudaMemcpyHostToDevice, h->stream));
    MCK(cudaStreamSynchronize(h->stream));   // `host` goes out of scope
    const size_t nd = (size_t)V.ext[0] * V.ext[1] * V.ext[2];
    if (S.dense_doubles < nd) { if (S.dense) MCK(cudaFree(S.dense)); MCK(cudaMalloc(&S.dense, nd * sizeof(double))); S.dense_doubles = nd; }
    V.dense = S.dense;
    for (int q = 0; q < 3; ++q) { V.dense_box.lo[q] = 0; V.dense_box.hi[q] = V.ext[q] - 1; }
    V.dense_box.ncomp = 1;
}
dim3 mac_mf_grid(const b200np_mfab* m)
{
    long long mx = 1;
    for (int f = 0; f < m->nfabs; ++f)
        mx = std::max(mx, (long long)(m->box[f].hi[0] - m->box[f].lo[0] + 1) * (m->box[f].hi[1] - m->box[f].lo[1] + 1) * (m->box[f].hi[2] - m->box[f].lo[2] + 1));
    return dim3((unsigned)std::min<long long>((mx + 255) / 256, 64), (unsigned)m->nfabs);
}
void mac_mf_gather(b200mac* h, MacMfView& V)
{
    MLAUNCH(h, k_mac_mf_gather, mac_mf_grid(V.m), 256, (const MacMfFab*)h->mf[V.slot].tab, V.m->ngrow, V.dense, V.ext[0], V.ext[1], V.ext[2]);
}
void mac_mf_scatter(b200mac* h, MacMfView& V, b200np_stats* st)
{
    MLAUNCH(h, k_mac_mf_scatter, mac_mf_grid(V.m), 256, (const MacMfFab*)h->mf[V.slot].tab, V.m->ngrow, (const double*)V.dense, V.ext[0], V.ext[1], V.ext[2]);
    if (!V.staged) return;
    size_t total = 0;
    for (int f = 0; f < V.m->nfabs; ++f) {
        const b200np_fab& b = V.m->box[f];
        const size_t nd = (size_t)(b.hi[0] - b.lo[0] + 1) * (b.hi[1] - b.lo[1] + 1) * (b.hi[2] - b.lo[2] + 1);
        MCK(cudaMemcpyAsync(V.m->data[f], h->mf[V.slot].stage + V.off[f], nd * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        total += nd;
    }
    st->d2h_bytes += (long long)(total * sizeof(double));
}
}  // namespace

extern "C" {

// initProjector / updateCoeffs over face-centred MultiFabs (inv_rho[lev][d], :71-90)
int b200mac_set_coeffs_mf(b200mac_t* h, const b200np_mfab* bx, const b200np_mfab* by, const b200np_mfab* bz)
{
    if (!h) return B200NP_ERR_BAD_ARG;
    const int* n = h->lv[0].g.n;
    const b200np_mfab* m[3] = {bx, by, bz};
    for (int d = 0; d < 3; ++d) if (!mac_mf_ok(m[d], n, d)) return B200NP_ERR_BAD_ARG;
    try {
        MCK(cudaSetDevice(h->device));
        b200np_stats st{};
        MacMfView V[3];
        for (int d = 0; d < 3; ++d) {
            mac_mf_map(h, V[d], d, m[d], n, d, true, &st);
            mac_mf_gather(h, V[d]);
        }
        MCK(cudaStreamSynchronize(h->stream));
        return b200mac_set_coeffs(h, V[0].dense, &V[0].dense_box, V[1].dense, &V[1].dense_box, V[2].dense, &V[2].dense_box, 0.0);
    } catch (int e) { return e; }
}

// project over multi-box MultiFabs: u_mac / v_mac / w_mac face-centred, mac_phi cell-centred (optional).  Results are bit-identical to
// the single-box call; only the valid faces / cells of every fab are written.
int b200mac_project_mf(b200mac_t* h, const b200np_mfab* umac, const b200np_mfab* vmac, const b200np_mfab* wmac, const b200np_mfab* mac_phi,
                       int phi_is_initial_guess, double rtol, double atol, b200np_stats* stats)
{
    b200np_stats local{};
    b200np_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof(*st));
    if (!h || !h->have_coeffs) return st->status = B200NP_ERR_BAD_ARG;
    const int* n = h->lv[0].g.n;
    const b200np_mfab* m[3] = {umac, vmac, wmac};
    for (int d = 0; d < 3; ++d) if (!mac_mf_ok(m[d], n, d)) return st->status = B200NP_ERR_BAD_ARG;
    if (mac_phi && !mac_mf_ok(mac_phi, n, -1)) return st->status = B200NP_ERR_BAD_ARG;
    try {
        MCK(cudaSetDevice(h->device));
        MacMfView V[4];
        for (int d = 0; d < 3; ++d) {
            mac_mf_map(h, V[d], 3 + d, m[d], n, d, true, st);
            mac_mf_gather(h, V[d]);
        }
        if (mac_phi) {
            mac_mf_map(h, V[3], 6, mac_phi, n, -1, phi_is_initial_guess != 0, st);
            if (phi_is_initial_guess) mac_mf_gather(h, V[3]);
        }
        MCK(cudaStreamSynchronize(h->stream));
        b200np_stats inner{};
        const int rc = b200mac_project(h, V[0].dense, &V[0].dense_box, V[1].dense, &V[1].dense_box, V[2].dense, &V[2].dense_box,
                                       mac_phi ? V[3].dense : nullptr, mac_phi ? &V[3].dense_box : nullptr, phi_is_initial_guess, rtol, atol, &inner);
        const long long h2d = st->h2d_bytes;
        *st = inner;
        st->h2d_bytes += h2d;
        for (int d = 0; d < 3; ++d) mac_mf_scatter(h, V[d], st);
        if (mac_phi) mac_mf_scatter(h, V[3], st);
        MCK(cudaStreamSynchronize(h->stream));
        return st->status = rc;
    } catch (int e) { return st->status = e; }
}

// test hooks: one building block on device-resident level arrays (dense cell layout, host arrays in / out)
int b200mac_level_op(b200mac_t* h, int lev, int op, int arg, const double* in_a, const double* in_b, double* out)
{
    if (!h || lev < 0 || lev >= (int)h->lv.size()) return B200NP_ERR_BAD_ARG;
    try {
        MCK(cudaSetDevice(h->device));
        MacLevel& L = h->lv[lev];
        const size_t bytes = L.ncell * sizeof(double);
        auto up = [&](double* d, const double* src) { if (src) MCK(cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, h->stream)); };
        switch (op) {
        case 0:   // smooth: cor <- arg smooth calls on (cor = in_a, res = in_b)
            up(L.cor, in_a); up(L.res, in_b);
            mac_smooth(h, L, L.cor, L.res, arg);
            MCK(cudaMemcpyAsync(out, L.cor, bytes, cudaMemcpyDeviceToHost, h->stream));
            break;
        case 1:   // residual: out = in_b - A in_a
            up(L.cor, in_a); up(L.res, in_b);
            MLAUNCH(h, k_mac_residual, mac_grid3(L.g), dim3(64, 4), L.g, (const double*)L.cor, (const double*)L.res, L.rescor, (double*)nullptr);
            MCK(cudaMemcpyAsync(out, L.rescor, bytes, cudaMemcpyDeviceToHost, h->stream));
            break;
        case 2: { // restrict: out (level lev+1) = R in_a
            if (lev + 1 >= (int)h->lv.size()) return B200NP_ERR_BAD_ARG;
            MacLevel& C = h->lv[lev + 1];
            up(L.rescor, in_a);
            MLAUNCH(h, k_mac_restrict, grid_for(C.ncell), 256, C.g, L.g.n[0], L.g.n[1], (const double*)L.rescor, C.res);
            MCK(cudaMemcpyAsync(out, C.res, C.ncell * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            break;
        }
        case 3: { // interpolate: out = in_a + P in_b (in_b on level lev+1)
            if (lev + 1 >= (int)h->lv.size()) return B200NP_ERR_BAD_ARG;
            MacLevel& C = h->lv[lev + 1];
            up(L.cor, in_a);
            if (in_b) MCK(cudaMemcpyAsync(C.cor, in_b, C.ncell * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            MLAUNCH(h, k_mac_interp_add, grid_for(L.ncell), 256, L.g, C.g.n[0], C.g.n[1], L.cor, (const double*)C.cor);
            MCK(cudaMemcpyAsync(out, L.cor, bytes, cudaMemcpyDeviceToHost, h->stream));
            break;
        }
        case 4: { // bottom solve on the coarsest level: out = solve(in_b)
            MacLevel& B = h->lv.back();
            if (in_b) MCK(cudaMemcpyAsync(B.res, in_b, B.ncell * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            MCK(cudaMemsetAsync(h->dinfo, 0, 4 * sizeof(int), h->stream));
            MLAUNCH(h, k_mac_bottom, 1, 1024, B.g, B.cor, B.res, h->work, h->opts.bottom_maxiter, h->opts.bottom_rtol, h->opts.bottom_atol,
                    h->singular ? 1 : 0, h->dinfo);
            MCK(cudaMemcpyAsync(out, B.cor, B.ncell * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            break;
        }
        default: return B200NP_ERR_BAD_ARG;
        }
        MCK(cudaStreamSynchronize(h->stream));
        return B200NP_OK;
    } catch (int e) { return e; }
}

int b200mac_level_dims(const b200mac_t* h, int lev, int n_cell[3])
{
    if (!h || lev < 0 || lev >= (int)h->lv.size()) return B200NP_ERR_BAD_ARG;
    for (int d = 0; d < 3; ++d) n_cell[d] = h->lv[lev].g.n[d];
    return B200NP_OK;
}

}  // extern "C"
