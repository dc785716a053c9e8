// b200np.cu -- host side of the B200-native nodal projection: multigrid hierarchy, MLMG
// driver, slab decomposition over the GPUs of one node (NCCL), C ABI (include/b200np.h).
// Host code is C++ and only launches the sm_100a kernels in np_kernels.cuh / np_smooth.cuh;
// there is no CPU compute path.
//
// Reference functions restated here (see include/b200np.h and DESIGN.md for the map):
//   incflo::ApplyNodalProjection      src/projection/incflo_apply_nodal_projection.cpp:29-267
//   Hydro::NodalProjector::project    [U] SURVEY.md A.1
//   MLMG::solve / mgVcycle            [U] SURVEY.md A.9
//   MLNodeLinOp::defineGrids, applyBC [U] SURVEY.md A.8 (hierarchy, FillBoundary)
//   FabArray::FillBoundary / ParallelAllReduce [U] -> NCCL send/recv of z planes, ncclAllReduce
#include "../../include/b200np.h"
#include "np_kernels.cuh"
#include "np_smooth.cuh"
#include "np_smooth3.cuh"
#include "np_composite.cuh"
#include "np_multibox.cuh"
#include "np_peer.h"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace b200np_dev;

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            fprintf(stderr, "b200np: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            throw int(B200NP_ERR_CUDA);                                                                  \
        }                                                                                                \
    } while (0)

namespace {

// NCCL is loaded lazily (dlopen) so that single-GPU users never depend on it.
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load()
    {
        if (lib) return true;
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return false;
#define NP_SYM(name) *(void**)(&name) = dlsym(lib, "nccl" #name); if (!name) return false;
        NP_SYM(GetUniqueId) NP_SYM(CommInitRank) NP_SYM(CommDestroy) NP_SYM(Send) NP_SYM(Recv) NP_SYM(AllReduce) NP_SYM(AllGather)
        NP_SYM(GroupStart) NP_SYM(GroupEnd) NP_SYM(GetErrorString)
#undef NP_SYM
        return true;
    }
};
NcclApi g_nccl;

#define NK(call)                                                                                       \
    do {                                                                                               \
        ncclResult_t r_ = (call);                                                                      \
        if (r_ != ncclSuccess) {                                                                       \
            fprintf(stderr, "b200np: NCCL error %s at %s:%d\n", g_nccl.GetErrorString(r_), __FILE__, __LINE__); \
            throw int(B200NP_ERR_NCCL);                                                                \
        }                                                                                              \
    } while (0)

b200np_peer::DriverApi g_drv;
std::atomic<unsigned long long> g_peer_serial{0};

// Every device array of a handle lives in ONE allocation.  Sizes are rank-independent, so an array
// sits at the same offset on every rank and a neighbour's copy is (peer base + my offset) once the
// neighbour's arena is mapped with CUDA IPC (slab-decomposed path).  Pass 1 measures, pass 2 places.
struct Arena {
    char* base = nullptr;
    size_t size = 0, off = 0;
    bool measure = true;
    double* take(size_t n_doubles)
    {
        const size_t b = (n_doubles * sizeof(double) + 1023) & ~size_t(1023);
        double* p = measure ? nullptr : reinterpret_cast<double*>(base + off);
        off += b;
        return p;
    }
};

struct LevelData {
    Lev g{};
    bool dist = false;        // slab-distributed level (ghost plane slots are exchanged)
    bool iso = false;         // dx == dy == dz: face coefficients of the stencil vanish (np_smooth3.cuh)
    double* sigma = nullptr;  // plane 0 of owned cells (allocation starts one plane earlier)
    double* sigma_alloc = nullptr;
    int nzl_alloc = 0;        // node planes allocated per array (rank-independent: the largest slab)
    double *sol = nullptr, *rhs = nullptr, *res = nullptr, *cor = nullptr, *cor2 = nullptr, *rescor = nullptr;
    dim3 gn, gc;     // grids of 64x4-thread blocks over owned nodes / cells
    dim3 gsm;        // smoother / residual grid (tiles x z-chunks)
    dim3 git;        // interpolation grid (fine tiles)
    int tz = 16;     // smoother z-chunk of this level
    long long nblk_n = 0;
    // first replicated level only: this rank's share as produced by restriction / sigma coarsening
    Lev gpart{};
    dim3 gn_part, gc_part;
    double *part_nodal = nullptr, *part_sigma = nullptr;
};

}  // namespace

struct b200np {
    b200np_geom geom{};
    b200np_opts opts{};
    int device = 0;
    int rank = 0, nranks = 1;
    ncclComm_t comm = nullptr;
    Arena arena;
    // NVLink peer-memory halos (CUDA IPC): the neighbours' arenas mapped into this process
    int use_p2p = 1;          // B200NP_P2P=0: grouped ncclSend/ncclRecv per halo instead
    int fuse_halo = 1;        // B200NP_FUSE_HALO=0: separate k_halo_pull before every sweep
    bool p2p = false;         // active (every rank mapped its neighbours)
    char *peer_lo = nullptr, *peer_hi = nullptr;
    // how the neighbours' arenas are mapped: 0 none, 1 cuMem* allocation + POSIX fd (np_peer.h), 2 legacy CUDA IPC
    int peer_map = 0;
    int peer_map_want = 1;    // B200NP_PEER_MAP=vmm|ipc (default vmm: legacy IPC refuses >= 1 GB arenas on some boxes)
    b200np_peer::VmmMapping arena_vmm{}, peer_lo_vmm{}, peer_hi_vmm{};
    unsigned long long* flags = nullptr;  // HaloFlags::my
    unsigned long long xk = 0;            // exchanges issued since the epoch base was last advanced
    double* ipc_buf = nullptr;
    int nlev_dist = 0;       // levels [0, nlev_dist) are slab-distributed, the rest replicated on every rank
    int singular = 1;
    bool var_sigma = false;
    std::vector<LevelData> lv;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    cudaEvent_t ev[6]{};
    double* partial = nullptr;  // reduction partials
    double* dscal = nullptr;    // device scalars: [0..1] sums, [2] norm
    double* hscal = nullptr;    // pinned host mirror
    int* dinfo = nullptr;       // bottom solver info (iters, ret)
    int* hinfo = nullptr;
    double* bottom_work = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    bool graph_direct = false;   // the captured V-cycle relaxes (sol, rhs) on level 0 (vcycle_launch's direct form)
    int top_direct = 1;          // B200NP_TOP_DIRECT=0: MLMG's correction form on level 0 too (cor = 0, smooth, sol += cor)
    bool graph_var = false;
    double graph_csig = 0.0;  // kernel parameters (Lev by value) are baked into the captured graph
    long long launches = 0, launches_per_vcycle = 0, exchanges = 0;
    struct Stage { double* d = nullptr; size_t bytes = 0; };
    Stage stage[8];  // staging buffers for host-pointer callers
    // multi-box callers (b200np_*_mf): per field the slab array the boxes are gathered into, the device table of fab
    // descriptors and, for host pointers, the staging area of the fabs
    struct MfSlot { double* slab = nullptr; size_t slab_bytes = 0; MfFab* tab = nullptr; size_t tab_cap = 0; double* stage = nullptr; size_t stage_bytes = 0; };
    MfSlot mf[6];
    cudaEvent_t mf_ev[2] = {nullptr, nullptr};
    int TZ = 64;
    int dist_graph = 1;       // capture the slab-decomposed V-cycle (NCCL send/recv included) into a CUDA graph (B200NP_DIST_GRAPH)
    int dist_min_planes = 64; // a level stays slab-distributed while every rank keeps at least this many cell planes (B200NP_DIST_MIN_PLANES;
                              // measured at 2 GPUs, 256^3 per GPU: 8 -> 33.1 ms, 64 -> 30.6 ms per solve)
    int use_pdl = 1;          // programmatic dependent launch between the V-cycle kernels (B200NP_PDL)
    int res_max_ctas = 148;   // levels with at most this many smoother CTAs use the resident-chunk kernel (B200NP_RES_CTAS)
    int smoother_version = 3, resid_version = 3;  // B200NP_SMOOTHER=2 / B200NP_RESID=2: the general (anisotropic) kernels on isotropic levels too
    // B200NP_PROFILE=1: per-phase device times (CUDA events between phases, graph capture off); printed per call
    int profile = 0;
    int zero_start = -1;      // skip the memset of cor before a pre-smooth and the read of it in the first sweep (-1.6 % per solve);
                              // default: on for one GPU, off on slabs (not yet measured there); B200NP_ZERO_START=0|1 overrides
    int interp_tz = 4;        // B200NP_INTERP_TZ = 8 | 4: fine planes per interpolation tile (4: 35 KB of shared memory,
                              // 5-6 CTAs per SM; measured 7-12 % faster than 8)
    int dbg_halo = 0;         // B200NP_DBG_HALO: see smooth_sweeps (timing experiments, results are wrong)
    bool has_profile = false; // b200np_set_inflow_profile: IncfloVelFill evaluated on the device
    InflowProfile profile_data{};
    int face_type[6] = {0, 0, 0, 0, 0, 0};   // b200np_set_face_types (enum b200np_face_type), amrex::Orientation order
    bool has_dd = false;                      // some face is direction_dependent: enforceInOutSolvability applies
    double inout_flux[2] = {0.0, 0.0};        // influx / outflux of the last call (b200np_inout_flux)
    bool no_bottom = false;   // fine AMR level of a composite solve: one MG level, never a bottom solve
    std::vector<cudaEvent_t> prof_ev;
    std::vector<std::string> prof_tag;
    size_t prof_n = 0;
};

namespace {

#define LAUNCH(h, kern, grid, block, ...)                        \
    do {                                                         \
        kern<<<grid, block, 0, (h)->stream>>>(__VA_ARGS__);      \
        (h)->launches++;                                         \
    } while (0)

// Launch with programmatic dependent launch (PDL): the kernel may become resident while its
// predecessor drains; every kernel launched this way calls pdl_wait() before its first access to
// data the predecessor produced (np_level.h).  Captured into the V-cycle graph as programmatic edges.
template <typename... KArgs, typename... Args>
void launch_pdl(b200np* h, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = h->use_pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, kern, KArgs(args)...));
    h->launches++;
}

// profiling: the time between two marks is charged to the tag of the later one
void prof_mark(b200np* h, const char* tag, int lev = -1)
{
    if (!h->profile) return;
    if (h->prof_n == h->prof_ev.size()) { cudaEvent_t e; CK(cudaEventCreate(&e)); h->prof_ev.push_back(e); h->prof_tag.emplace_back(); }
    h->prof_tag[h->prof_n] = lev >= 0 ? std::string("L") + std::to_string(lev) + " " + tag : std::string(tag);
    CK(cudaEventRecord(h->prof_ev[h->prof_n++], h->stream));
}
void prof_report(b200np* h)
{
    if (!h->profile || h->prof_n < 2) { h->prof_n = 0; return; }
    CK(cudaEventSynchronize(h->prof_ev[h->prof_n - 1]));
    std::map<std::string, std::pair<double, int>> acc;
    double tot = 0;
    for (size_t i = 1; i < h->prof_n; ++i) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, h->prof_ev[i - 1], h->prof_ev[i]));
        auto& a = acc[h->prof_tag[i]];
        a.first += ms; a.second++;
        tot += ms;
    }
    fprintf(stderr, "b200np profile (rank %d): %.3f ms between first and last mark\n", h->rank, tot);
    for (auto& kv : acc)
        fprintf(stderr, "  %-22s n=%5d %9.3f ms %5.1f %%  avg %8.1f us\n", kv.first.c_str(), kv.second.second, kv.second.first,
                100.0 * kv.second.first / tot, 1e3 * kv.second.first / kv.second.second);
    h->prof_n = 0;
}

inline double* dev_alloc(b200np* h, size_t n_doubles) { return h->arena.take(n_doubles); }

// allocate a nodal array with ghost plane slots; returns pointer to owned plane 0
double* alloc_nodal(b200np* h, LevelData& L)
{
    double* base = dev_alloc(h, (size_t)L.g.ps * (L.nzl_alloc + 2));
    return base + L.g.ps;
}

inline bool zper(const b200np* h) { return h->geom.bc_lo[2] == B200NP_BC_PERIODIC; }

// fills the level-independent part of a descriptor for global cell counts n and spacing dx
void fill_lev(const b200np_geom& G, const int n[3], const double dx[3], Lev& g)
{
    for (int d = 0; d < 3; ++d) {
        g.n[d] = n[d];
        g.per[d] = (G.bc_lo[d] == B200NP_BC_PERIODIC);
        g.nn[d] = n[d] + (g.per[d] ? 0 : 1);
        g.rlo[d] = G.bc_lo[d] == B200NP_BC_NEUMANN ? 1 : (G.bc_lo[d] == B200NP_BC_INFLOW ? 2 : 0);
        g.rhi[d] = G.bc_hi[d] == B200NP_BC_NEUMANN ? 1 : (G.bc_hi[d] == B200NP_BC_INFLOW ? 2 : 0);
        g.dlo[d] = G.bc_lo[d] == B200NP_BC_DIRICHLET;
        g.dhi[d] = G.bc_hi[d] == B200NP_BC_DIRICHLET;
        g.dxinv[d] = 1.0 / dx[d];
    }
    g.px = (g.nn[0] + 7) / 8 * 8;
    g.ps = (long long)g.px * g.nn[1];
    g.cpx = (g.n[0] + 7) / 8 * 8;
    g.cps = (long long)g.cpx * g.n[1];
    g.k0 = 0; g.nzl = g.nn[2];
    g.ck0 = 0; g.cnzl = g.n[2];
    g.dist = 0;
    const double fx = g.dxinv[0] * g.dxinv[0] / 36.0, fy = g.dxinv[1] * g.dxinv[1] / 36.0, fz = g.dxinv[2] * g.dxinv[2] / 36.0;
    g.fxyz = fx + fy + fz;
    g.fmx2y2z = -fx + 2 * fy + 2 * fz; g.f2xmy2z = 2 * fx - fy + 2 * fz; g.f2x2ymz = 2 * fx + 2 * fy - fz;
    g.f4xm2ym2z = 4 * fx - 2 * fy - 2 * fz; g.fm2x4ym2z = -2 * fx + 4 * fy - 2 * fz; g.fm2xm2y4z = -2 * fx - 2 * fy + 4 * fz;
    g.csig = 1.0; g.sigma = nullptr;
}

// this rank's slab of a level with n2 cell planes split over P ranks
void set_slab(Lev& g, int rank, int P, bool periodic_z)
{
    const int m = g.n[2] / P;
    g.ck0 = rank * m; g.cnzl = m;
    g.k0 = rank * m; g.nzl = m + ((!periodic_z && rank == P - 1) ? 1 : 0);
    g.dist = 1;
}

// ---- hierarchy plan (pure host logic; also exported as b200np_dist_plan for the CPU tests) ----
// level 0 is always distributed; a coarser level stays distributed while every rank keeps an even
// number (>= min_planes) of cell planes, or while it is too big to replicate (> 128^3 nodes)
// (a replicated level costs every rank a sweep over ALL its nodes, so a big level with thin slabs --
// 512^3 on 8 GPUs: level 1 = 256^3 with 32 planes per rank -- stays distributed down to 8 planes)
bool level_stays_dist(const int n[3], int P, int lev, int min_planes)
{
    if (n[2] % P != 0 || (n[2] / P) % 2 != 0) return false;
    if (lev == 0) return true;
    const double gnodes = (double)(n[0] + 1) * (n[1] + 1) * (n[2] + 1);
    return n[2] / P >= min_planes || (gnodes > 2.2e6 && n[2] / P >= 8);
}
// coarsen by 2 while every direction stays even and >= 2 cells wide (A.8)
bool can_coarsen(const int n[3], int lev_next, int max_coarsening_level)
{
    if (lev_next > max_coarsening_level || lev_next >= 30) return false;
    for (int d = 0; d < 3; ++d) if (n[d] % 2 != 0 || n[d] / 2 < 2) return false;
    return true;
}

void build_levels(b200np* h)
{
    const b200np_geom& G = h->geom;
    h->lv.clear();
    const int P = h->nranks;
    int n[3] = {G.n_cell[0], G.n_cell[1], G.n_cell[2]};
    double dx[3] = {G.dx[0], G.dx[1], G.dx[2]};
    h->singular = 1;
    for (int d = 0; d < 3; ++d)
        if (G.bc_lo[d] == B200NP_BC_DIRICHLET || G.bc_hi[d] == B200NP_BC_DIRICHLET) h->singular = 0;
    int lev = 0;
    bool still_dist = P > 1;
    h->nlev_dist = 0;
    for (;;) {
        LevelData L;
        Lev& g = L.g;
        fill_lev(G, n, dx, g);
        L.iso = (dx[0] == dx[1] && dx[1] == dx[2]);
        if (still_dist && level_stays_dist(n, P, lev, h->dist_min_planes)) {
            set_slab(g, h->rank, P, zper(h));
            L.dist = true;
            h->nlev_dist = lev + 1;
        } else {
            if (still_dist && lev == 0) throw int(B200NP_ERR_BAD_ARG);  // level 0 must be distributable
            if (still_dist) {  // first replicated level: remember this rank's share
                L.gpart = g;
                set_slab(L.gpart, h->rank, P, zper(h));
                L.gn_part = dim3((g.nn[0] + 63) / 64, (g.nn[1] + 3) / 4, L.gpart.nzl);
                L.gc_part = dim3((g.n[0] + 63) / 64, (g.n[1] + 3) / 4, L.gpart.cnzl);
                L.part_nodal = dev_alloc(h, (size_t)g.ps * (n[2] / P + 1 + 2));
                L.part_sigma = dev_alloc(h, (size_t)g.cps * (L.gpart.cnzl + 2));
            }
            still_dist = false;
        }
        L.gn = dim3((g.nn[0] + 63) / 64, (g.nn[1] + 3) / 4, g.nzl);
        L.gc = dim3((g.n[0] + 63) / 64, (g.n[1] + 3) / 4, g.cnzl);
        // z-chunk rule (mirrored by the oracle): aim at 592 CTAs per sweep = 148 SMs x 2 resident
        // CTAs x 2 waves, chunk height between 8 and the cap opts.tile[2] (< 8 costs a V-cycle)
        {
            const int ntiles = ((g.nn[0] + NP_TX - 1) / NP_TX) * ((g.nn[1] + NP_TY - 1) / NP_TY);
            const int nch = std::max(1, 592 / ntiles);
            // lower bound: 8 planes on levels with more than 33 node planes (shorter chunks cost a V-cycle);
            // coarse levels are insensitive and use 4 / 2 / 1 so that more chunks run in parallel
            const int mn = g.nn[2] > 33 ? 8 : g.nn[2] > 17 ? 4 : g.nn[2] > 9 ? 2 : 1;
            // slab-decomposed level: the chunks tile this rank's planes (rank-independent count), so that a
            // sweep still fills the SMs twice; the converged answer does not depend on the chunking
            const int nzp = L.dist ? n[2] / P + (g.per[2] ? 0 : 1) : g.nn[2];
            L.tz = std::max(mn, std::min(h->TZ, (nzp + nch - 1) / nch));
        }
        L.gsm = dim3((g.nn[0] + NP_TX - 1) / NP_TX, (g.nn[1] + NP_TY - 1) / NP_TY, (g.nzl + L.tz - 1) / L.tz);
        L.git = dim3((g.nn[0] + IT_X - 1) / IT_X, (g.nn[1] + IT_Y - 1) / IT_Y, (g.nzl + h->interp_tz - 1) / h->interp_tz);
        L.nblk_n = (long long)L.gn.x * L.gn.y * L.gn.z;
        L.nzl_alloc = L.dist ? n[2] / P + 1 : g.nzl;   // the last slab of a non-periodic domain owns one more plane
        L.sigma_alloc = dev_alloc(h, (size_t)g.cps * (g.cnzl + 2));
        L.sigma = L.sigma_alloc + g.cps;
        L.res = alloc_nodal(h, L); L.cor = alloc_nodal(h, L); L.cor2 = alloc_nodal(h, L); L.rescor = alloc_nodal(h, L);
        if (lev == 0) { L.sol = alloc_nodal(h, L); L.rhs = alloc_nodal(h, L); }
        h->lv.push_back(L);
        ++lev;
        if (!can_coarsen(n, lev, h->opts.mg_max_coarsening_level)) break;
        for (int d = 0; d < 3; ++d) { n[d] /= 2; dx[d] *= 2; }
    }
    if (P > 1 && h->nlev_dist >= (int)h->lv.size()) throw int(B200NP_ERR_BAD_ARG);  // needs a replicated coarse level
    long long maxblk = 0;  // rank-independent bound on the per-block reduction partials
    for (auto& L : h->lv) maxblk = std::max(maxblk, (long long)L.gn.x * L.gn.y * L.nzl_alloc);
    h->partial = dev_alloc(h, (size_t)2 * maxblk + 16);
    h->dscal = dev_alloc(h, 16);
    h->flags = reinterpret_cast<unsigned long long*>(dev_alloc(h, 16));
    h->ipc_buf = dev_alloc(h, (size_t)8 * std::max(P, 1) + 8);   // 64 bytes per rank
    h->dinfo = reinterpret_cast<int*>(dev_alloc(h, 8));
    const Lev& B = h->lv.back().g;
    h->bottom_work = dev_alloc(h, h->no_bottom ? 8 : (size_t)B.ps * B.nzl * 8);
}

void build_hierarchy(b200np* h)
{
    h->arena = Arena{};
    build_levels(h);                       // pass 1: sizes
    h->arena.size = h->arena.off;
    void* base = nullptr;
    // slab handles: an exportable cuMem* allocation, so the neighbours can map it over a POSIX fd (setup_p2p)
    if (h->nranks > 1 && h->use_p2p && h->peer_map_want == 1) {
        std::string why;
        if (g_drv.load() && b200np_peer::vmm_alloc(g_drv, h->device, h->arena.size, h->arena_vmm, &why)) base = (void*)h->arena_vmm.ptr;
        else fprintf(stderr, "b200np[rank %d]: cuMem* arena unavailable (%s), using cudaMalloc + legacy CUDA IPC\n", h->rank, why.c_str());
        cudaGetLastError();
    }
    if (!base) CK(cudaMalloc(&base, h->arena.size));
    // The memset runs on the legacy default stream, which the handle's non-blocking stream does not wait for: without the
    // synchronisation the first small copies on h->stream (the records / IPC handles setup_p2p gathers through this very
    // arena) can land BEFORE a large arena has been cleared and are then wiped -- the "invalid argument" from
    // cudaIpcOpenMemHandle seen on 4-GPU boxes in round 1 was a zeroed handle, not an IPC limitation.
    CK(cudaMemset(base, 0, h->arena.size));
    CK(cudaDeviceSynchronize());
    h->arena.base = static_cast<char*>(base);
    h->arena.off = 0; h->arena.measure = false;
    build_levels(h);                       // pass 2: pointers
    CK(cudaMallocHost(&h->hscal, 16 * sizeof(double)));
    CK(cudaMallocHost(&h->hinfo, 16 * sizeof(int)));
}

// ---- slab communication (MLNodeLinOp::applyBC's FillBoundary) ------------------------
// Exchanges one plane with each z neighbour: plane `first` -> lower neighbour's upper ghost slot,
// plane `last` -> upper neighbour's lower ghost slot.  Physical (non-periodic) ends are filled
// locally by `end_lo` / `end_hi` (reflection for nodes, clamp for cells).
// peer copy of one of my arena arrays
template <typename T>
inline T* peer_ptr(const b200np* h, const char* peer_base, T* mine)
{
    return reinterpret_cast<T*>(const_cast<char*>(peer_base) + (reinterpret_cast<const char*>(mine) - h->arena.base));
}
// flags of the next exchange (one call per exchange: it takes the next epoch number)
inline HaloFlags halo_flags(b200np* h)
{
    const int P = h->nranks, r = h->rank;
    const bool per = zper(h);
    HaloFlags f{};
    f.my = h->flags;
    if (per || r > 0) f.lo_flag = peer_ptr(h, h->peer_lo, h->flags) + 1;      // I am my lower neighbour's upper neighbour
    if (per || r < P - 1) f.hi_flag = peer_ptr(h, h->peer_hi, h->flags) + 0;
    f.k = h->xk++;
    return f;
}
// advance the device-side epoch base by the exchanges issued since the last advance (np_kernels.cuh K10)
void epoch_advance(b200np* h)
{
    if (!h->p2p || h->xk == 0) return;
    LAUNCH(h, k_epoch_advance, 1, 1, h->flags, h->xk);
    h->xk = 0;
}
// handshake without data: everything both neighbours launched before it is complete when it returns
void p2p_fence(b200np* h)
{
    if (!h->p2p) return;
    launch_pdl(h, k_halo_pull, dim3(1), dim3(256), 0, halo_flags(h), (double2*)nullptr, (const double2*)nullptr, (double2*)nullptr,
               (const double2*)nullptr, 0ll);
}
void exchange_planes(b200np* h, double* base, long long plane, int nown, int m, bool nodal)
{
    const int P = h->nranks, r = h->rank;
    const bool per = zper(h);
    const bool has_lo = per || r > 0, has_hi = per || r < P - 1;
    const int lo = (r - 1 + P) % P, hi = (r + 1) % P;
    double* ghost_lo = base - plane;
    double* ghost_hi = base + (long long)nown * plane;
    // physical (non-periodic) ends -- nodes: phi(-1) = phi(1); cells: copy of the adjacent interior cell
    const double* end_lo = base + (nodal ? plane : 0);
    const double* end_hi = base + (long long)(nown - (nodal ? 2 : 1)) * plane;
    h->exchanges++;
    if (h->p2p) {
        // pull the lower neighbour's top owned plane (every rank but the last owns m planes; the lower
        // neighbour is the last rank only through the periodic wrap) and the upper neighbour's plane 0.
        // WAR safety (np_kernels.cuh K10): callers never overwrite an exchanged array before the next
        // exchange -- sweeps ping-pong, residual / restriction / interpolation write other arrays.
        const double* src_lo = has_lo ? peer_ptr(h, h->peer_lo, base) + (long long)(m - 1) * plane : end_lo;
        const double* src_hi = has_hi ? peer_ptr(h, h->peer_hi, base) : end_hi;
        const long long n2 = plane / 2;
        const int nb = (int)std::max(1ll, std::min(64ll, (n2 + 511) / 512));
        launch_pdl(h, k_halo_pull, dim3(nb), dim3(256), 0, halo_flags(h), (double2*)ghost_lo, (const double2*)src_lo, (double2*)ghost_hi,
                   (const double2*)src_hi, n2);
        return;
    }
    NK(g_nccl.GroupStart());
    if (has_lo) NK(g_nccl.Send(base, plane, ncclDouble, lo, h->comm, h->stream));
    if (has_hi) NK(g_nccl.Send(base + (long long)(nown - 1) * plane, plane, ncclDouble, hi, h->comm, h->stream));
    if (has_hi) NK(g_nccl.Recv(ghost_hi, plane, ncclDouble, hi, h->comm, h->stream));
    if (has_lo) NK(g_nccl.Recv(ghost_lo, plane, ncclDouble, lo, h->comm, h->stream));
    NK(g_nccl.GroupEnd());
    if (!has_lo) CK(cudaMemcpyAsync(ghost_lo, end_lo, plane * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if (!has_hi) CK(cudaMemcpyAsync(ghost_hi, end_hi, plane * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
}
// Map the neighbours' arenas.  Preferred: the arena is a cuMem* allocation whose POSIX file descriptor goes to
// the neighbours over a UNIX socket (np_peer.h); else legacy CUDA IPC handles exchanged with ncclAllGather.
// Falls back -- loudly -- to NCCL send/recv halos on every rank if any rank cannot map its neighbours.
void setup_p2p(b200np* h)
{
    const int P = h->nranks, r = h->rank;
    if (P == 1 || !h->use_p2p) return;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    const bool per = zper(h);
    const bool has_lo = per || r > 0, has_hi = per || r < P - 1;
    const int lo = (r - 1 + P) % P, hi = (r + 1) % P;
    char* buf = reinterpret_cast<char*>(h->ipc_buf);
    double ok = 1.0;
    auto all_min = [&](double v) {
        CK(cudaMemcpyAsync(h->dscal + 4, &v, sizeof(double), cudaMemcpyHostToDevice, h->stream));
        NK(g_nccl.AllReduce(h->dscal + 4, h->dscal + 4, 1, ncclDouble, ncclMin, h->comm, h->stream));
        CK(cudaMemcpyAsync(&v, h->dscal + 4, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return v;
    };
    // ---- (1) cuMem* arena + POSIX fd ----
    struct Rec { long long pid; unsigned long long serial, size; long long vmm; char pad[32]; };
    static_assert(sizeof(Rec) == 64, "record size");
    std::vector<Rec> recs(P);
    Rec me{};
    me.pid = (long long)getpid(); me.serial = g_peer_serial.fetch_add(1); me.size = h->arena_vmm.size; me.vmm = h->arena_vmm.mapped ? 1 : 0;
    int sock = -1;
    if (me.vmm) {
        sock = b200np_peer::fd_socket_bind(me.pid, me.serial);
        if (sock < 0) { fprintf(stderr, "b200np[rank %d]: UNIX socket for the arena descriptor: %s\n", r, strerror(errno)); me.vmm = 0; }
    }
    CK(cudaMemcpyAsync(buf + 64 * r, &me, 64, cudaMemcpyHostToDevice, h->stream));
    NK(g_nccl.AllGather(buf + 64 * r, buf, 64, ncclChar, h->comm, h->stream));
    CK(cudaMemcpyAsync(recs.data(), buf, (size_t)64 * P, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    bool all_vmm = true;
    for (const Rec& q : recs) all_vmm = all_vmm && q.vmm;
    if (all_vmm) {
        std::string why;
        int fd = b200np_peer::vmm_export_fd(g_drv, h->arena_vmm, &why);
        if (fd < 0) { fprintf(stderr, "b200np[rank %d]: %s\n", r, why.c_str()); ok = 0.0; }
        int expect = 0;
        auto send_to = [&](int peer) {
            if (fd >= 0 && !b200np_peer::fd_send(sock, recs[peer].pid, recs[peer].serial, fd, r)) {
                fprintf(stderr, "b200np[rank %d]: sending the arena descriptor to rank %d: %s\n", r, peer, strerror(errno));
                ok = 0.0;
            }
            ++expect;
        };
        // every rank sends even after a local failure would leave a neighbour waiting: recv has a 20 s timeout
        if (has_lo) send_to(lo);
        if (has_hi && !(has_lo && hi == lo)) send_to(hi);
        for (int i = 0; i < expect; ++i) {
            int from = -1;
            const int pfd = b200np_peer::fd_recv(sock, &from);
            if (pfd < 0 || from < 0 || from >= P) {
                fprintf(stderr, "b200np[rank %d]: no arena descriptor from a neighbour (%s)\n", r, strerror(errno));
                ok = 0.0;
                if (pfd >= 0) close(pfd);
                continue;
            }
            b200np_peer::VmmMapping m;
            if (!b200np_peer::vmm_import(g_drv, h->device, pfd, recs[from].size, m, &why)) {
                fprintf(stderr, "b200np[rank %d]: mapping the arena of rank %d (%llu bytes): %s\n", r, from, recs[from].size, why.c_str());
                ok = 0.0;
            } else if (from == lo && has_lo && !h->peer_lo_vmm.mapped) h->peer_lo_vmm = m;
            else if (from == hi && has_hi && !h->peer_hi_vmm.mapped) h->peer_hi_vmm = m;
            else b200np_peer::vmm_free(g_drv, m);
            close(pfd);
        }
        if (fd >= 0) close(fd);
        cudaGetLastError();
        if (has_lo && !h->peer_lo_vmm.mapped) ok = 0.0;
        if (has_hi && !(has_lo && hi == lo) && !h->peer_hi_vmm.mapped) ok = 0.0;
        ok = all_min(ok);
        if (sock >= 0) close(sock);
        if (ok > 0) {
            h->peer_lo = has_lo ? (char*)h->peer_lo_vmm.ptr : nullptr;
            h->peer_hi = !has_hi ? nullptr : (has_lo && hi == lo) ? h->peer_lo : (char*)h->peer_hi_vmm.ptr;
            h->p2p = true; h->peer_map = 1;
            return;
        }
        b200np_peer::vmm_free(g_drv, h->peer_lo_vmm);
        b200np_peer::vmm_free(g_drv, h->peer_hi_vmm);
    } else {
        if (sock >= 0) close(sock);
        // ---- (2) legacy CUDA IPC (only when no rank holds a cuMem* arena: those cannot be exported this way) ----
        bool none_vmm = true;
        for (const Rec& q : recs) none_vmm = none_vmm && !q.vmm && q.size == 0;
        std::vector<cudaIpcMemHandle_t> all(P);
        cudaIpcMemHandle_t mine{};
        if (!none_vmm) ok = 0.0;
        else {
            cudaError_t e = cudaIpcGetMemHandle(&mine, h->arena.base);
            if (e != cudaSuccess) {
                fprintf(stderr, "b200np[rank %d]: cudaIpcGetMemHandle failed: %s\n", r, cudaGetErrorString(e));
                ok = 0.0;
            }
            cudaGetLastError();
        }
        CK(cudaMemcpyAsync(buf + 64 * r, &mine, 64, cudaMemcpyHostToDevice, h->stream));
        NK(g_nccl.AllGather(buf + 64 * r, buf, 64, ncclChar, h->comm, h->stream));
        CK(cudaMemcpyAsync(all.data(), buf, (size_t)64 * P, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        void *plo = nullptr, *phi = nullptr;
        auto open_peer = [&](void** p, int peer) {
            cudaError_t e = cudaIpcOpenMemHandle(p, all[peer], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                fprintf(stderr, "b200np[rank %d]: cudaIpcOpenMemHandle(arena of rank %d, %zu bytes) failed: %s\n", r, peer, h->arena.size,
                        cudaGetErrorString(e));
                *p = nullptr;
                ok = 0.0;
            }
            cudaGetLastError();
        };
        if (ok > 0 && has_lo) open_peer(&plo, lo);
        if (ok > 0 && has_hi) {
            if (has_lo && hi == lo) phi = plo;
            else open_peer(&phi, hi);
        }
        ok = all_min(ok);
        if (ok > 0) {
            h->peer_lo = static_cast<char*>(plo); h->peer_hi = static_cast<char*>(phi);
            h->p2p = true; h->peer_map = 2;
            return;
        }
        if (plo) cudaIpcCloseMemHandle(plo);
        if (phi && phi != plo) cudaIpcCloseMemHandle(phi);
        cudaGetLastError();
    }
    h->peer_lo = h->peer_hi = nullptr;
    h->p2p = false; h->peer_map = 0;
    if (r == 0) fprintf(stderr, "b200np: WARNING: peer mapping of the neighbours' arenas unavailable on at least one rank, slab halos FALL BACK to "
                                "ncclSend/ncclRecv (b200np_halo_transport() == 2)\n");
}
inline void halo_nodes(b200np* h, LevelData& L, double* x)
{
    if (L.dist) exchange_planes(h, x, L.g.ps, L.g.nzl, L.g.cnzl, true);
}
inline void halo_cells(b200np* h, LevelData& L, double* s)
{
    if (L.dist) exchange_planes(h, s, L.g.cps, L.g.cnzl, L.g.cnzl, false);
}
// vel.FillBoundary(1 ghost) across slabs, written into the caller's ghost cell planes (compRHS, A.2)
void halo_vel(b200np* h, Fab vel)
{
    if (h->nranks == 1) return;
    const Lev& g = h->lv[0].g;
    const int P = h->nranks, r = h->rank;
    const bool per = zper(h);
    const bool has_lo = per || r > 0, has_hi = per || r < P - 1;
    const int lo = (r - 1 + P) % P, hi = (r + 1) % P;
    const long long pl = (long long)vel.nx * vel.ny;
    const int zfirst = g.ck0 - vel.lo[2], zlast = g.ck0 + g.cnzl - 1 - vel.lo[2];
    NK(g_nccl.GroupStart());
    for (int c = 0; c < 3; ++c) {
        double* b = vel.p + c * vel.cstride;
        if (has_lo) NK(g_nccl.Send(b + zfirst * pl, pl, ncclDouble, lo, h->comm, h->stream));
        if (has_hi) NK(g_nccl.Send(b + zlast * pl, pl, ncclDouble, hi, h->comm, h->stream));
        if (has_hi) NK(g_nccl.Recv(b + (zlast + 1) * pl, pl, ncclDouble, hi, h->comm, h->stream));
        if (has_lo) NK(g_nccl.Recv(b + (zfirst - 1) * pl, pl, ncclDouble, lo, h->comm, h->stream));
    }
    NK(g_nccl.GroupEnd());
    h->exchanges++;
}
// assemble a replicated array from every rank's share (agglomeration onto all ranks):
// rank q's planes [q*m, q*m + cnt_q) of `full` come from q's `part`.
void allgather_parts(b200np* h, const Lev& gfull, const double* part, double* full, bool nodal)
{
    const int P = h->nranks, r = h->rank;
    const long long plane = nodal ? gfull.ps : gfull.cps;
    const int m = gfull.n[2] / P;
    auto cnt = [&](int q) { return m + ((nodal && !zper(h) && q == P - 1) ? 1 : 0); };
    NK(g_nccl.GroupStart());
    for (int q = 0; q < P; ++q) {
        if (q == r) continue;
        NK(g_nccl.Send(part, (size_t)cnt(r) * plane, ncclDouble, q, h->comm, h->stream));
        NK(g_nccl.Recv(full + (long long)q * m * plane, (size_t)cnt(q) * plane, ncclDouble, q, h->comm, h->stream));
    }
    NK(g_nccl.GroupEnd());
    CK(cudaMemcpyAsync(full + (long long)r * m * plane, part, (size_t)cnt(r) * plane * sizeof(double),
                       cudaMemcpyDeviceToDevice, h->stream));
    h->exchanges++;
}
inline void allreduce(b200np* h, double* d, int n, ncclRedOp_t op)
{
    if (h->nranks > 1) NK(g_nccl.AllReduce(d, d, n, ncclDouble, op, h->comm, h->stream));
}

// ---- multigrid building blocks --------------------------------------------------------------------
void set_level_sigma_ptrs(b200np* h, bool var, double csig)
{
    h->var_sigma = var;
    for (auto& L : h->lv) {
        L.g.sigma = var ? L.sigma : nullptr; L.g.csig = csig;
        L.gpart.sigma = nullptr; L.gpart.csig = csig;
    }
}

// average_down of sigma to every level (A.8) + ghost layers across slabs
void coarsen_sigma(b200np* h)
{
    if (!h->var_sigma) return;
    halo_cells(h, h->lv[0], h->lv[0].sigma);
    for (size_t l = 0; l + 1 < h->lv.size(); ++l) {
        LevelData &F = h->lv[l], &C = h->lv[l + 1];
        if (F.dist && !C.dist) {  // first replicated level: coarsen my share, then gather everybody's
            LAUNCH(h, k_coarsen_sigma, C.gc_part, 256, F.g, C.gpart, F.sigma, C.part_sigma + C.g.cps);
            allgather_parts(h, C.g, C.part_sigma + C.g.cps, C.sigma, false);
        } else {
            LAUNCH(h, k_coarsen_sigma, C.gc, 256, F.g, C.g, F.sigma, C.sigma);
            halo_cells(h, C, C.sigma);
        }
    }
}

// Gauss-Seidel sweeps (ping-pong x -> y, then swap); halo refresh before every sweep on slab levels
void smooth_sweeps(b200np* h, LevelData& L, double*& x, double*& y, const double* rhs, int nsweeps, bool zero_start = false)
{
    const bool iso_path = h->smoother_version >= 3 && L.iso;
    const bool resident = L.tz <= SM_RES_TZ && (int)(L.gsm.x * L.gsm.y * L.gsm.z) <= h->res_max_ctas;
    // slab level over NVLink peer memory: the halo exchange is part of the sweep kernel (HaloFused)
    const bool fused = L.dist && h->p2p && h->fuse_halo && iso_path;
    for (int s = 0; s < nsweeps; ++s) {
        // first sweep of a zero-start smooth call: the kernel does not read x at all (and nobody zeroed it)
        const int tzarg = L.tz | ((zero_start && s == 0 && h->zero_start && iso_path) ? SM_ZERO_IN : 0);
        if (fused) {
            const int P = h->nranks, r = h->rank;
            const bool per = zper(h);
            const bool has_lo = per || r > 0, has_hi = per || r < P - 1;
            const Lev& g = L.g;
            // first sweep: the input halo comes from a standalone exchange (nothing to do if x == 0 everywhere)
            if (s == 0 && !zero_start && h->dbg_halo < 2) halo_nodes(h, L, x);
            HaloFused H{};
            H.f = halo_flags(h);
            H.pin_lo = has_lo ? x - g.ps : x + g.ps;                                    // ghost slot / reflection plane
            H.pin_hi = has_hi ? x + (long long)g.nzl * g.ps : x + (long long)(g.nzl - 2) * g.ps;
            // the lower neighbour owns cnzl planes (it is the last rank only through the periodic wrap)
            H.out_lo = has_lo ? peer_ptr(h, h->peer_lo, y) + (long long)g.cnzl * g.ps : nullptr;
            H.out_hi = has_hi ? peer_ptr(h, h->peer_hi, y) - g.ps : nullptr;
            // same chunks as the unfused sweep; the top one (the remainder) is scheduled first
            const int nch = (g.nzl + L.tz - 1) / L.tz;
            H.tztop = g.nzl - (nch - 1) * L.tz;
            H.first = s == 0; H.more = s + 1 < nsweeps;
            // timing experiments only (wrong halos): 1 no flags, 2 + no remote stores, 3 + no epoch ticket,
            // 4 no flags, no ticket, remote stores kept; 5 (unfused path) no exchange at all
            if (h->dbg_halo >= 1) { H.first = 1; H.more = 0; }
            if (h->dbg_halo == 2 || h->dbg_halo == 3) { H.out_lo = nullptr; H.out_hi = nullptr; }
            if (h->dbg_halo == 3 || h->dbg_halo == 4) H.first = 2;
            if (s == 0 && h->dbg_halo >= 1) {}
            h->exchanges++;
            const dim3 grid(L.gsm.x, L.gsm.y, nch);
            if (resident) {
                if (h->var_sigma) launch_pdl(h, k_smooth_iso_res_dist<true>, grid, dim3(256), SM_RES_DOUBLES * sizeof(double), g, x, y, rhs, tzarg, H);
                else              launch_pdl(h, k_smooth_iso_res_dist<false>, grid, dim3(256), (SM_RES_TZ + 2) * SM_PHI_SLOT * sizeof(double), g, x, y, rhs, tzarg, H);
            } else {
                if (h->var_sigma) launch_pdl(h, k_smooth_iso_dist<true>, grid, dim3(256), SM_SMOOTH_DOUBLES * sizeof(double), g, x, y, rhs, tzarg, H);
                else              launch_pdl(h, k_smooth_iso_dist<false>, grid, dim3(256), 4 * SM_PHI_SLOT * sizeof(double), g, x, y, rhs, tzarg, H);
            }
            std::swap(x, y);
            continue;
        }
        if (h->dbg_halo != 5) halo_nodes(h, L, x);
        if (!iso_path) {
            if (h->var_sigma) launch_pdl(h, k_smooth_v2<true>, L.gsm, dim3(256), SM_SMOOTH_DOUBLES * sizeof(double), L.g, x, y, rhs, L.tz);
            else              launch_pdl(h, k_smooth_v2<false>, L.gsm, dim3(256), 4 * SM_PHI_SLOT * sizeof(double), L.g, x, y, rhs, L.tz);
        } else if (resident) {
            // small isotropic level: whole chunk resident in shared memory, one CTA per SM
            if (h->var_sigma) launch_pdl(h, k_smooth_iso_res<true>, L.gsm, dim3(256), SM_RES_DOUBLES * sizeof(double), L.g, x, y, rhs, tzarg);
            else              launch_pdl(h, k_smooth_iso_res<false>, L.gsm, dim3(256), (SM_RES_TZ + 2) * SM_PHI_SLOT * sizeof(double), L.g, x, y, rhs, tzarg);
        } else {  // isotropic level: 2-barrier / register-carried variant, same semantics
            if (h->var_sigma) launch_pdl(h, k_smooth_iso<true>, L.gsm, dim3(256), SM_SMOOTH_DOUBLES * sizeof(double), L.g, x, y, rhs, tzarg);
            else              launch_pdl(h, k_smooth_iso<false>, L.gsm, dim3(256), 4 * SM_PHI_SLOT * sizeof(double), L.g, x, y, rhs, tzarg);
        }
        std::swap(x, y);
    }
}

// res = rhs - L phi (phi's ghost planes are refreshed first on slab levels).
// gov / var_override: evaluate with another descriptor of the same arrays (other boundary conditions or
// another sigma array) -- the composite solver's one-sided sums (np_composite.cuh).
void residual(b200np* h, LevelData& L, double* phi, const double* rhs, double* res, double* norm_partial,
              const Lev* gov = nullptr, int var_override = -1)
{
    halo_nodes(h, L, phi);
    const Lev& g = gov ? *gov : L.g;
    const bool var = var_override >= 0 ? var_override != 0 : h->var_sigma;
    if (h->resid_version == 2 || !L.iso) {
        if (var) launch_pdl(h, k_residual_v2<true>, L.gsm, dim3(256), SM_SMOOTH_DOUBLES * sizeof(double), g, phi, rhs, res, L.tz, norm_partial);
        else     launch_pdl(h, k_residual_v2<false>, L.gsm, dim3(256), 4 * SM_PHI_SLOT * sizeof(double), g, phi, rhs, res, L.tz, norm_partial);
    } else {
        if (var) launch_pdl(h, k_residual_iso<true>, L.gsm, dim3(256), SM_SMOOTH_DOUBLES * sizeof(double), g, phi, rhs, res, L.tz, norm_partial);
        else     launch_pdl(h, k_residual_iso<false>, L.gsm, dim3(256), 4 * SM_PHI_SLOT * sizeof(double), g, phi, rhs, res, L.tz, norm_partial);
    }
}
// number of per-CTA norm partials the residual kernel writes
long long resid_nblk(b200np* h, LevelData& L)
{
    return (long long)L.gsm.x * L.gsm.y * L.gsm.z;
}

void bottom_solve(b200np* h)
{
    LevelData& B = h->lv.back();
    if (h->var_sigma)
        launch_pdl(h, k_bottom_bicgstab<true>, dim3(1), dim3(512), 0, B.g, B.cor, B.res, h->bottom_work, h->opts.bottom_maxiter,
               h->opts.bottom_rtol, h->opts.bottom_atol, h->singular, h->opts.smooth_num_sweeps, h->opts.bottom_solver, h->dinfo);
    else
        launch_pdl(h, k_bottom_bicgstab<false>, dim3(1), dim3(512), 0, B.g, B.cor, B.res, h->bottom_work, h->opts.bottom_maxiter,
               h->opts.bottom_rtol, h->opts.bottom_atol, h->singular, h->opts.smooth_num_sweeps, h->opts.bottom_solver, h->dinfo);
}

void restrict_to(b200np* h, int l)
{
    LevelData &F = h->lv[l], &C = h->lv[l + 1];
    halo_nodes(h, F, F.rescor);
    if (F.dist && !C.dist) {  // agglomeration: my share of the coarse rhs, then gather onto every rank
        launch_pdl(h, k_restrict, C.gn_part, dim3(256), 0, F.g, C.gpart, (const double*)F.rescor, C.part_nodal + C.g.ps);
        allgather_parts(h, C.g, C.part_nodal + C.g.ps, C.res, true);
    } else {
        launch_pdl(h, k_restrict, C.gn, dim3(256), 0, F.g, C.g, (const double*)F.rescor, C.res);
    }
}
void interp_add(b200np* h, int l, double* fine = nullptr)
{
    LevelData &F = h->lv[l], &C = h->lv[l + 1];
    if (!fine) fine = F.cor;
    halo_nodes(h, C, C.cor);
    {
        if (h->interp_tz == 4) {
            if (h->var_sigma) launch_pdl(h, k_interp_tile<true, 4>, F.git, dim3(256), (it_v_doubles(4) + it_s_doubles(4)) * sizeof(double), F.g, C.g, fine, (const double*)C.cor);
            else              launch_pdl(h, k_interp_tile<false, 4>, F.git, dim3(256), it_v_doubles(4) * sizeof(double), F.g, C.g, fine, (const double*)C.cor);
        } else {
            if (h->var_sigma) launch_pdl(h, k_interp_tile<true>, F.git, dim3(256), (IT_V_DOUBLES + IT_S_DOUBLES) * sizeof(double), F.g, C.g, fine, (const double*)C.cor);
            else              launch_pdl(h, k_interp_tile<false>, F.git, dim3(256), IT_V_DOUBLES * sizeof(double), F.g, C.g, fine, (const double*)C.cor);
        }
    }
}

// MLMG::mgVcycle (A.9) on (cor, res), all launches on h->stream, no host synchronisation.
// direct (level 0 of mlmg_solve only): the cycle relaxes (sol, rhs) in place of (cor, res) on the finest level.  The
// smoother is a stationary linear iteration, so smoothing cor from 0 against res = rhs - L sol and adding it to sol is
// the same arithmetic as smoothing sol against rhs (rounding aside); the direct form needs neither the zeroed cor,
// nor `sol += cor` (24 B/node), nor the stored top residual (8 B/node) -- MLMG::oneIter's result is unchanged.
void vcycle_launch(b200np* h, int lev0, bool direct = false)
{
    const int nsw = h->opts.smooth_num_sweeps;
    const int nl = (int)h->lv.size();
    for (int l = lev0; l < nl - 1; ++l) {
        LevelData& L = h->lv[l];
        const bool dtop = direct && l == 0;
        double* X = dtop ? L.sol : L.cor;
        const double* R = dtop ? L.rhs : L.res;
        // cor = 0 (ghost slots included) -- unless the first sweep knows it and never reads cor
        const bool skip_zero = dtop || (h->zero_start && h->smoother_version >= 3 && L.iso && h->opts.num_pre_smooth * nsw >= 1);
        if (!skip_zero) CK(cudaMemsetAsync(L.cor - L.g.ps, 0, (size_t)L.g.ps * (L.g.nzl + 2) * sizeof(double), h->stream));
        double *x = X, *y = L.cor2;
        prof_mark(h, "zero cor", l);
        smooth_sweeps(h, L, x, y, R, h->opts.num_pre_smooth * nsw, !dtop);  // cor == 0, ghost slots included
        if (x != X) {  // odd sweep count: cor was pulled by the neighbours in the last exchange
            if (L.dist) p2p_fence(h);
            CK(cudaMemcpyAsync(X, x, (size_t)L.g.ps * L.g.nzl * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        }
        prof_mark(h, "smooth", l);
        residual(h, L, X, R, L.rescor, nullptr);
        prof_mark(h, "residual", l);
        restrict_to(h, l);
        prof_mark(h, "restrict", l);
    }
    bottom_solve(h);
    prof_mark(h, "bottom");
    for (int l = nl - 2; l >= lev0; --l) {
        LevelData& L = h->lv[l];
        const bool dtop = direct && l == 0;
        double* X = dtop ? L.sol : L.cor;
        interp_add(h, l, X);
        prof_mark(h, "interp", l);
        double *x = X, *y = L.cor2;
        smooth_sweeps(h, L, x, y, dtop ? L.rhs : L.res, h->opts.num_post_smooth * nsw);
        if (x != X) {
            if (L.dist) p2p_fence(h);
            CK(cudaMemcpyAsync(X, x, (size_t)L.g.ps * L.g.nzl * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        }
        prof_mark(h, "smooth", l);
    }
}

void vcycle(b200np* h, bool direct = false)
{
    if (!h->opts.use_graph || h->profile || (h->nranks > 1 && !h->dist_graph)) { vcycle_launch(h, 0, direct); return; }
    if (h->graph_exec && (h->graph_var != h->var_sigma || h->graph_csig != h->lv[0].g.csig || h->graph_direct != direct)) {
        cudaGraphExecDestroy(h->graph_exec); cudaGraphDestroy(h->graph);
        h->graph_exec = nullptr; h->graph = nullptr;
    }
    epoch_advance(h);   // the graph's exchanges are numbered from 0
    if (!h->graph_exec) {
        long long before = h->launches;
        CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        try {
            vcycle_launch(h, 0, direct);
            epoch_advance(h);   // last node: every replay leaves the base advanced by the graph's exchanges
        } catch (int) {   // never leave the (possibly caller-owned) stream in capture mode
            cudaGraph_t broken = nullptr;
            cudaStreamEndCapture(h->stream, &broken);
            if (broken) cudaGraphDestroy(broken);
            cudaGetLastError();
            throw;
        }
        CK(cudaStreamEndCapture(h->stream, &h->graph));
        CK(cudaGraphInstantiate(&h->graph_exec, h->graph, 0));
        h->launches_per_vcycle = h->launches - before;
        h->launches = before;
        h->graph_var = h->var_sigma;
        h->graph_direct = direct;
        h->graph_csig = h->lv[0].g.csig;
    }
    CK(cudaGraphLaunch(h->graph_exec, h->stream));
    h->launches += h->launches_per_vcycle;
}

// global inf-norm from per-CTA partials: device max, ncclAllReduce(max) over slabs, one host read
double norm_from_partials(b200np* h, long long nb)
{
    LAUNCH(h, k_max_final, 1, 1024, h->partial, nb, h->dscal + 2);
    allreduce(h, h->dscal + 2, 1, ncclMax);
    CK(cudaMemcpyAsync(h->hscal + 2, h->dscal + 2, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return h->hscal[2];
}

double norminf(b200np* h, LevelData& L, const double* x)
{
    LAUNCH(h, k_norminf_partial, L.gn, 256, L.g, x, h->partial);
    return norm_from_partials(h, L.nblk_n);
}

// MLMG::solve (A.9) on level-0 arrays sol (initial guess, in/out) and rhs
// sol_is_zero: the initial guess is identically zero (NodalProjector::project always starts there), so the initial residual
// IS rhs -- no kernel needed when the cycle never reads L0.res (direct form)
int mlmg_solve(b200np* h, double rtol, double atol, b200np_stats* st, bool sol_is_zero = false)
{
    LevelData& L0 = h->lv[0];
    st->iters = 0; st->bottom_iters = 0; st->status = B200NP_OK; st->nlevels = (int)h->lv.size();
    CK(cudaMemsetAsync(h->dinfo, 0, 4 * sizeof(int), h->stream));
    if (!h->singular) {
        if (!sol_is_zero) LAUNCH(h, k_zero_masked, L0.gn, 256, L0.g, L0.sol);
        LAUNCH(h, k_zero_masked, L0.gn, 256, L0.g, L0.rhs);
        st->rhsnorm = norminf(h, L0, L0.rhs);
    } else {  // makeSolvable: subtract the weighted mean of rhs (A.8); the inf-norm of the result comes out of the same pass
        LAUNCH(h, k_wsum_partial, L0.gn, 256, L0.g, L0.rhs, h->partial);
        LAUNCH(h, k_sum2_final, 1, 1024, h->partial, L0.nblk_n, h->dscal);
        allreduce(h, h->dscal, 2, ncclSum);
        LAUNCH(h, k_sub_mean, L0.gn, 256, L0.g, L0.rhs, (const double*)h->dscal, h->partial);
        st->rhsnorm = norm_from_partials(h, L0.nblk_n);
    }
    const bool direct0 = sol_is_zero && h->top_direct && h->lv.size() > 1;
    if (direct0) st->resnorm0 = st->rhsnorm;
    else {
        residual(h, L0, L0.sol, L0.rhs, L0.res, h->partial);
        st->resnorm0 = norm_from_partials(h, resid_nblk(h, L0));
    }
    const double maxnorm = std::max(st->rhsnorm, st->resnorm0);
    const double target = std::max(atol, std::max(rtol, 1e-16) * maxnorm);
    st->resnorm = st->resnorm0;
    st->resnorm_hist[0] = st->resnorm0;
    const bool talk = h->rank == 0;
    if (talk && h->opts.verbose >= 1) printf("MLMG: Initial rhs               = %.12g\nMLMG: Initial residual (resid0) = %.12g\n", st->rhsnorm, st->resnorm0);
    if (st->resnorm0 <= target) return B200NP_OK;
    bool converged = false;
    prof_mark(h, "mlmg setup");
    for (int it = 0; it < h->opts.maxiter; ++it) {
        const bool direct = h->top_direct && h->lv.size() > 1;
        vcycle(h, direct);
        if (!direct) LAUNCH(h, k_axpy, L0.gn, 256, L0.g, L0.sol, L0.cor, 1.0);
        // direct form: only the norm of the new residual is needed (the next cycle never reads res)
        residual(h, L0, L0.sol, L0.rhs, direct ? nullptr : L0.res, h->partial);
        prof_mark(h, direct ? "top residual norm" : "top sol+=cor, residual");
        st->resnorm = norm_from_partials(h, resid_nblk(h, L0));
        prof_mark(h, "top norm (allreduce+sync)");
        st->iters = it + 1;
        if (it + 1 < 128) st->resnorm_hist[it + 1] = st->resnorm;
        if (talk && h->opts.verbose >= 2) printf("MLMG: Iteration %3d Fine resid/bnorm = %.12g\n", it + 1, st->resnorm / maxnorm);
        if (st->resnorm <= target) { converged = true; break; }
        if (!(st->resnorm <= 1e20 * maxnorm)) { st->status = B200NP_ERR_DIVERGED; break; }
    }
    if (!converged && st->status == B200NP_OK) st->status = B200NP_ERR_NOT_CONVERGED;
    CK(cudaMemcpyAsync(h->hinfo, h->dinfo, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    if (h->p2p) CK(cudaMemcpyAsync(h->hscal + 8, h->flags + 6, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    st->bottom_iters = h->hinfo[0];
    if (h->p2p && reinterpret_cast<unsigned long long*>(h->hscal + 8)[0] != 0ull) {   // halo_spin gave up on a neighbour
        CK(cudaMemsetAsync(h->flags + 6, 0, sizeof(unsigned long long), h->stream));
        st->status = B200NP_ERR_PEER_TIMEOUT;
    }
    if (talk && h->opts.verbose >= 1)
        printf("MLMG: Final Iter. %d resid, resid/bnorm = %.12g, %.12g\n", st->iters, st->resnorm, st->resnorm / maxnorm);
    return st->status;
}

// ---- caller arrays ------------------------------------------------------------------------
Fab make_fab(double* p, const b200np_fab* b)
{
    Fab f{};
    f.p = p;
    if (!b) return f;
    for (int d = 0; d < 3; ++d) f.lo[d] = b->lo[d];
    f.nx = b->hi[0] - b->lo[0] + 1; f.ny = b->hi[1] - b->lo[1] + 1; f.nz = b->hi[2] - b->lo[2] + 1;
    f.cstride = (long long)f.nx * f.ny * f.nz;
    return f;
}
size_t fab_bytes(const b200np_fab* b)
{
    return (size_t)(b->hi[0] - b->lo[0] + 1) * (b->hi[1] - b->lo[1] + 1) * (b->hi[2] - b->lo[2] + 1) * b->ncomp * sizeof(double);
}
bool is_device_ptr(const void* p)
{
    if (!p) return true;
    cudaPointerAttributes a{};
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
// returns a device pointer for `p` (staging host memory when needed)
double* stage_in(b200np* h, int slot, const double* p, const b200np_fab* box, bool copy, bool* staged, b200np_stats* st)
{
    *staged = false;
    if (!p) return nullptr;
    if (is_device_ptr(p)) return const_cast<double*>(p);
    size_t bytes = fab_bytes(box);
    auto& S = h->stage[slot];
    if (S.bytes < bytes) {
        if (S.d) CK(cudaFree(S.d));
        CK(cudaMalloc(&S.d, bytes));
        S.bytes = bytes;
    }
    if (copy) {
        CK(cudaMemcpyAsync(S.d, p, bytes, cudaMemcpyHostToDevice, h->stream));
        st->h2d_bytes += (long long)bytes;
    }
    *staged = true;
    return S.d;
}
void stage_out(b200np* h, int slot, double* p, const b200np_fab* box, bool staged, b200np_stats* st)
{
    if (!staged || !p) return;
    size_t bytes = fab_bytes(box);
    CK(cudaMemcpyAsync(p, h->stage[slot].d, bytes, cudaMemcpyDeviceToHost, h->stream));
    st->d2h_bytes += (long long)bytes;
}

bool box_covers(const b200np_fab* b, const int lo[3], const int hi[3], int ncomp)
{
    if (!b || b->ncomp < ncomp) return false;
    for (int d = 0; d < 3; ++d) if (b->lo[d] > lo[d] || b->hi[d] < hi[d]) return false;
    return true;
}
// the caller's velocity box must hold one ghost layer wherever the divergence reads it:
// at non-periodic faces and, on a slab, towards the z neighbours
bool vel_box_ok(const b200np* h, const b200np_fab* vb)
{
    const Lev& g = h->lv[0].g;
    const int clo[3] = {0, 0, g.ck0}, chi[3] = {g.n[0] - 1, g.n[1] - 1, g.ck0 + g.cnzl - 1};
    if (!box_covers(vb, clo, chi, 3)) return false;
    for (int d = 0; d < 3; ++d) {
        const bool need = !g.per[d] || (d == 2 && h->nranks > 1);
        if (need && (vb->lo[d] > clo[d] - 1 || vb->hi[d] < chi[d] + 1)) return false;
    }
    return true;
}

// the caller's nodal box (phi / p_nd) must hold every node plane this rank writes: all nodes [0, n] in x and y and, in z,
// the rank's cell slab's nodes [ck0, ck0 + cnzl]; on a slab it must not reach beyond them (k_copy_phi reads the node
// plane above the slab from the ghost slot and nothing further away)
bool nodal_box_ok(const b200np* h, const b200np_fab* pb)
{
    const Lev& g = h->lv[0].g;
    const int nlo[3] = {0, 0, g.ck0}, nhi[3] = {g.n[0], g.n[1], g.ck0 + g.cnzl};
    if (!box_covers(pb, nlo, nhi, 1)) return false;
    if (h->nranks > 1 && (pb->lo[2] < nlo[2] - (g.ck0 > 0 || g.per[2] ? 1 : 0) || pb->hi[2] > nhi[2])) return false;
    return true;
}

// the common core: rhs = D vel; solve; vel -= sigma G phi; gphi, phi copy-out.
int project_core(b200np* h, Fab vel, Fab velo, int add_old, Fab gphi, int acc_g, Fab pout, int acc_p, double rtol,
                 double atol, b200np_stats* st)
{
    LevelData& L0 = h->lv[0];
    prof_mark(h, "start");
    coarsen_sigma(h);
    prof_mark(h, "coarsen sigma");
    halo_vel(h, vel);
    LAUNCH(h, k_divu, L0.gn, 256, L0.g, vel, L0.rhs);
    prof_mark(h, "halo vel + divu");
    CK(cudaMemsetAsync(L0.sol - L0.g.ps, 0, (size_t)L0.g.ps * (L0.g.nzl + 2) * sizeof(double), h->stream));
    CK(cudaEventRecord(h->ev[2], h->stream));
    int status = mlmg_solve(h, rtol, atol, st, true);
    CK(cudaEventRecord(h->ev[3], h->stream));
    halo_nodes(h, L0, L0.sol);  // the gradient and the copy-out read the node plane above the slab
    LAUNCH(h, k_mknewu, L0.gc, 256, L0.g, L0.sol, vel, velo, add_old, gphi, acc_g);
    if (pout.p) {
        dim3 g((pout.nx + 63) / 64, (pout.ny + 3) / 4, pout.nz);
        LAUNCH(h, k_copy_phi, g, 256, L0.g, L0.sol, pout, acc_p);
    }
    prof_mark(h, "mknewu + copy out");
    prof_report(h);
    return status;
}

void finish_stats(b200np* h, b200np_stats* st)
{
    CK(cudaEventRecord(h->ev[1], h->stream));
    CK(cudaEventSynchronize(h->ev[1]));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1])); st->ms_total = ms;
    CK(cudaEventElapsedTime(&ms, h->ev[2], h->ev[3])); st->ms_solve = ms;
    st->launches = h->launches;
}

int check_geom(const b200np_geom* g)
{
    for (int d = 0; d < 3; ++d) {
        if (g->n_cell[d] < 2 || !(g->dx[d] > 0)) return B200NP_ERR_BAD_ARG;
        if (g->bc_lo[d] < 0 || g->bc_lo[d] > 3 || g->bc_hi[d] < 0 || g->bc_hi[d] > 3) return B200NP_ERR_BAD_BC;
        if ((g->bc_lo[d] == B200NP_BC_PERIODIC) != (g->bc_hi[d] == B200NP_BC_PERIODIC)) return B200NP_ERR_BAD_BC;
    }
    return B200NP_OK;
}

int create_common(b200np_t** out, const b200np_geom* geom, const b200np_opts* opts, int device, int rank, int nranks,
                  const void* nccl_unique_id, bool fine_level = false)
{
    if (!out || !geom) return B200NP_ERR_BAD_ARG;
    *out = nullptr;
    int rc = check_geom(geom);
    if (rc) return rc;
    if (nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !nccl_unique_id)) return B200NP_ERR_BAD_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        return B200NP_ERR_CUDA;
    }
    if (nranks > 1 && !g_nccl.load()) return B200NP_ERR_NCCL;
    b200np* h = new b200np();
    try {
        CK(cudaSetDevice(device));
        h->device = device;
        h->no_bottom = fine_level;
        h->rank = rank; h->nranks = nranks;
        h->geom = *geom;
        if (opts) h->opts = *opts; else b200np_default_opts(&h->opts);
        if (h->opts.tile[0] != NP_TX || h->opts.tile[1] != NP_TY || h->opts.tile[2] < 1) { delete h; return B200NP_ERR_BAD_ARG; }
        h->TZ = h->opts.tile[2];
        if (const char* e = getenv("B200NP_SMOOTHER")) h->smoother_version = atoi(e);
        if (const char* e = getenv("B200NP_RES_CTAS")) h->res_max_ctas = atoi(e);
        if (const char* e = getenv("B200NP_DIST_GRAPH")) h->dist_graph = atoi(e);
        if (const char* e = getenv("B200NP_PDL")) h->use_pdl = atoi(e);
        if (const char* e = getenv("B200NP_PROFILE")) h->profile = atoi(e);
        if (const char* e = getenv("B200NP_DBG_HALO")) h->dbg_halo = atoi(e);
        if (const char* e = getenv("B200NP_P2P")) h->use_p2p = atoi(e);
        if (const char* e = getenv("B200NP_PEER_MAP")) h->peer_map_want = strcmp(e, "ipc") == 0 ? 2 : 1;
        if (const char* e = getenv("B200NP_FUSE_HALO")) h->fuse_halo = atoi(e);
        if (const char* e = getenv("B200NP_DIST_MIN_PLANES")) h->dist_min_planes = std::max(8, atoi(e));
        if (const char* e = getenv("B200NP_TOP_DIRECT")) h->top_direct = atoi(e);
        if (const char* e = getenv("B200NP_ZERO_START")) h->zero_start = atoi(e);
        if (const char* e = getenv("B200NP_INTERP_TZ")) h->interp_tz = atoi(e) == 8 ? 8 : 4;
        if (const char* e = getenv("B200NP_RESID")) h->resid_version = atoi(e);
        CK(cudaFuncSetAttribute(k_residual_v2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SM_SMOOTH_DOUBLES * sizeof(double))));
        CK(cudaFuncSetAttribute(k_residual_v2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * SM_PHI_SLOT * sizeof(double))));
        CK(cudaFuncSetAttribute(k_smooth_v2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SM_SMOOTH_DOUBLES * sizeof(double))));
        CK(cudaFuncSetAttribute(k_smooth_v2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * SM_PHI_SLOT * sizeof(double))));
        CK(cudaFuncSetAttribute(k_smooth_iso<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SM_SMOOTH_DOUBLES * sizeof(double))));
        CK(cudaFuncSetAttribute(k_smooth_iso<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * SM_PHI_SLOT * sizeof(double))));
        CK(cudaFuncSetAttribute(k_residual_iso<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SM_SMOOTH_DOUBLES * sizeof(double))));
        CK(cudaFuncSetAttribute(k_residual_iso<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * SM_PHI_SLOT * sizeof(double))));
        CK(cudaFuncSetAttribute(k_interp_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((IT_V_DOUBLES + IT_S_DOUBLES) * sizeof(double))));
        CK(cudaFuncSetAttribute(k_interp_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(IT_V_DOUBLES * sizeof(double))));
        CK(cudaFuncSetAttribute(k_interp_tile<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((it_v_doubles(4) + it_s_doubles(4)) * sizeof(double))));
        CK(cudaFuncSetAttribute(k_interp_tile<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(it_v_doubles(4) * sizeof(double))));
        CK(cudaFuncSetAttribute(k_smooth_iso_dist<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SM_SMOOTH_DOUBLES * sizeof(double))));
        CK(cudaFuncSetAttribute(k_smooth_iso_dist<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * SM_PHI_SLOT * sizeof(double))));
        CK(cudaFuncSetAttribute(k_smooth_iso_res_dist<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SM_RES_DOUBLES * sizeof(double))));
        CK(cudaFuncSetAttribute(k_smooth_iso_res_dist<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((SM_RES_TZ + 2) * SM_PHI_SLOT * sizeof(double))));
        CK(cudaFuncSetAttribute(k_smooth_iso_res<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SM_RES_DOUBLES * sizeof(double))));
        CK(cudaFuncSetAttribute(k_smooth_iso_res<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((SM_RES_TZ + 2) * SM_PHI_SLOT * sizeof(double))));
        CK(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
        h->stream = h->own_stream;
        for (auto& e : h->ev) CK(cudaEventCreate(&e));
        if (nranks > 1) {
            ncclUniqueId id;
            memcpy(&id, nccl_unique_id, sizeof(id));
            NK(g_nccl.CommInitRank(&h->comm, nranks, id, rank));
        }
        if (h->zero_start < 0) h->zero_start = nranks == 1 ? 1 : 0;
        build_hierarchy(h);
        setup_p2p(h);
        CK(cudaDeviceSynchronize());
    } catch (int e) {
        b200np_destroy(h);
        return e;
    }
    *out = h;
    return B200NP_OK;
}


// ============================================================================================
// Composite (two AMR level) projection: coarse level + ONE fine box at ratio 2 (BASELINE configs[3]).
// Algorithm and its justification: oracle/composite.py (MLMG::oneIter multi-level branch; with ratio 2
// the fine AMR level has a single MG level, so its miniCycle is nu1 smooth calls).
// ============================================================================================
}  // namespace

struct b200np_composite {
    b200np* h0 = nullptr;   // coarse level: the ordinary single-level hierarchy over the whole domain
    b200np* h1 = nullptr;   // fine box as a domain of its own with Dirichlet faces (coarse/fine interface nodes are not relaxed)
    CBox box{};
    int nb[3]{};            // covered coarse cells per direction
    Lev view{};             // the box nodes inside the coarse level's arrays, seen as a level (restriction target / interpolation source)
    long long off_n = 0;    // box node (lo) in a coarse nodal array
    double* rhsN = nullptr; // D u1 with reflecting box faces (2^f x one-sided divergence on the box boundary)
    double* s0z_alloc = nullptr;
    double* s0z = nullptr;  // coarse sigma with 0 in the covered cells
};

namespace {

// fine box with reflecting faces: one-sided sums x 2^f (np_composite.cuh)
Lev comp_gN(const b200np_composite* C)
{
    Lev g = C->h1->lv[0].g;   // interface faces: Dirichlet -> reflecting; wall / outflow / periodic faces stay what they are
    for (int d = 0; d < 3; ++d) {
        if (C->box.cf_lo[d]) { g.rlo[d] = 1; g.dlo[d] = 0; }
        if (C->box.cf_hi[d]) { g.rhi[d] = 1; g.dhi[d] = 0; }
    }
    return g;
}
dim3 box_grid(int nx, int ny, int nz) { return dim3((nx + 63) / 64, (ny + 3) / 4, nz); }

void comp_set_sigma(b200np_composite* C, bool var, double csig)
{
    b200np *h0 = C->h0, *h1 = C->h1;
    LevelData &L0 = h0->lv[0], &L1 = h1->lv[0];
    set_level_sigma_ptrs(h0, var, csig);
    set_level_sigma_ptrs(h1, var, csig);
    LAUNCH(h0, k_comp_sigma, L0.gc, 256, L0.g, L1.g, C->box, var ? L0.sigma : (double*)nullptr, C->s0z,
           var ? (const double*)L1.sigma : (const double*)nullptr, csig);
    coarsen_sigma(h0);   // averageDownCoeffs of the coarse hierarchy, with the fine average in the covered cells
}

// composite residual on the coarse level (MLNodeLaplacian::reflux): L0.res = rhs0 - A_composite(sol0, sol1)
void comp_coarse_residual(b200np_composite* C, const double* sol0, const double* sol1, const double* sums)
{
    b200np *h0 = C->h0, *h1 = C->h1;
    LevelData &L0 = h0->lv[0], &L1 = h1->lv[0];
    const Lev gN = comp_gN(C);
    residual(h1, L1, const_cast<double*>(sol1), C->rhsN, L1.rescor, nullptr, &gN);   // 2^f x (b - A) fine-side sums
    launch_pdl(h0, k_restrict, box_grid(C->box.nbn[0], C->box.nbn[1], C->box.nbn[2]), dim3(256), 0, gN, C->view,
               (const double*)L1.rescor, L0.rescor + C->off_n);
    Lev g0z = L0.g;
    g0z.sigma = C->s0z;
    residual(h0, L0, const_cast<double*>(sol0), L0.rhs, L0.res, nullptr, &g0z, 1);   // uncovered coarse cells only
    LAUNCH(h0, k_comp_combine, L0.gn, 256, L0.g, C->box, L0.res, (const double*)L0.rescor, sums);
}

double comp_read_norm(b200np_composite* C, long long nb1)
{
    b200np *h0 = C->h0, *h1 = C->h1;
    LevelData& L0 = h0->lv[0];
    LAUNCH(h0, k_comp_norm_excl, L0.gn, 256, L0.g, C->box, (const double*)L0.res, h0->partial);
    LAUNCH(h0, k_max_final, 1, 1024, h0->partial, L0.nblk_n, h0->dscal + 2);
    LAUNCH(h0, k_max_final, 1, 1024, h1->partial, nb1, h0->dscal + 3);
    CK(cudaMemcpyAsync(h0->hscal + 2, h0->dscal + 2, 2 * sizeof(double), cudaMemcpyDeviceToHost, h0->stream));
    CK(cudaStreamSynchronize(h0->stream));
    return std::max(h0->hscal[2], h0->hscal[3]);
}

// Hydro::NodalProjector::project over two levels.  Fabs of the fine level are indexed relative to the
// fine box (cell 0 = first fine cell of the box).
int comp_core(b200np_composite* C, Fab vel0, Fab vel1, Fab velo0, Fab velo1, int add_old, Fab gphi0, Fab gphi1, int acc_g,
              Fab p0, Fab p1, int acc_p, double rtol, double atol, b200np_stats* st)
{
    b200np *h0 = C->h0, *h1 = C->h1;
    LevelData &L0 = h0->lv[0], &L1 = h1->lv[0];
    const Lev gD = L1.g, gN = comp_gN(C);
    const CBox& b = C->box;
    const dim3 gbox = box_grid(C->nb[0], C->nb[1], C->nb[2]);
    const size_t bytes0 = (size_t)L0.g.ps * L0.g.nzl * sizeof(double), bytes1 = (size_t)L1.g.ps * L1.g.nzl * sizeof(double);
    const int nsw = h0->opts.smooth_num_sweeps;
    st->iters = 0; st->bottom_iters = 0; st->status = B200NP_OK; st->nlevels = (int)h0->lv.size() + 1;
    CK(cudaMemsetAsync(h0->dinfo, 0, 4 * sizeof(int), h0->stream));
    // ---- rhs (compRHS): covered coarse cells do not count; fine ghost cells are zero (:137) ----
    LAUNCH(h0, k_comp_zero_cells, gbox, 256, b, vel0, 3);
    LAUNCH(h0, k_divu, L0.gn, 256, L0.g, vel0, L0.rhs);          // uncovered coarse cells
    LAUNCH(h0, k_divu, L1.gn, 256, gD, vel1, L1.rhs);            // fine interior nodes (0 on the interface)
    LAUNCH(h0, k_divu, L1.gn, 256, gN, vel1, C->rhsN);           // + 2^f x one-sided sums on the interface
    CK(cudaMemsetAsync(L0.sol, 0, bytes0, h0->stream));
    CK(cudaMemsetAsync(L1.sol, 0, bytes1, h0->stream));
    CK(cudaMemsetAsync(L0.rescor, 0, bytes0, h0->stream));       // only its box nodes are ever written again
    CK(cudaEventRecord(h0->ev[2], h0->stream));
    // composite coarse rhs = residual of sol = 0: solvability offset (MLMG::makeSolvable, one offset for every level)
    double* sums = nullptr;
    comp_coarse_residual(C, L0.sol, L1.sol, nullptr);
    if (h0->singular) {
        LAUNCH(h0, k_wsum_partial, L0.gn, 256, L0.g, (const double*)L0.res, h0->partial);
        LAUNCH(h0, k_sum2_final, 1, 1024, h0->partial, L0.nblk_n, h0->dscal + 8);
        sums = h0->dscal + 8;
        LAUNCH(h0, k_sub_mean, L0.gn, 256, L0.g, L0.res, (const double*)sums);
        LAUNCH(h0, k_sub_mean, L1.gn, 256, gD, L1.rhs, (const double*)sums);
        LAUNCH(h0, k_zero_masked, L1.gn, 256, gD, L1.rhs);
    }
    residual(h1, L1, L1.sol, L1.rhs, L1.res, h1->partial);       // = rhs1 (sol1 = 0), with its norm
    const long long nb1 = resid_nblk(h1, L1);
    st->rhsnorm = st->resnorm0 = comp_read_norm(C, nb1);
    const double maxnorm = st->rhsnorm;
    const double target = std::max(atol, std::max(rtol, 1e-16) * maxnorm);
    st->resnorm = st->resnorm0;
    st->resnorm_hist[0] = st->resnorm0;
    const bool talk = h0->opts.verbose >= 1;
    if (talk) printf("MLMG: Initial rhs               = %.12g\nMLMG: Initial residual (resid0) = %.12g\n", st->rhsnorm, st->resnorm0);
    bool converged = st->resnorm0 <= target;
    for (int it = 0; !converged && it < h0->opts.maxiter; ++it) {
        // fine level: miniCycle = nu1 smooth calls on (cor, res), homogeneous Dirichlet on the interface
        CK(cudaMemsetAsync(L1.cor, 0, bytes1, h0->stream));
        double *x = L1.cor, *y = L1.cor2;
        smooth_sweeps(h1, L1, x, y, L1.res, h0->opts.num_pre_smooth * nsw, true);
        LAUNCH(h0, k_axpy, L1.gn, 256, gD, L1.sol, (const double*)x, 1.0);
        // coarse level: composite residual, solvability, V-cycle
        comp_coarse_residual(C, L0.sol, L1.sol, sums);
        if (h0->singular) {
            LAUNCH(h0, k_wsum_partial, L0.gn, 256, L0.g, (const double*)L0.res, h0->partial);
            LAUNCH(h0, k_sum2_final, 1, 1024, h0->partial, L0.nblk_n, h0->dscal);
            LAUNCH(h0, k_sub_mean, L0.gn, 256, L0.g, L0.res, (const double*)h0->dscal);
        }
        vcycle(h0);
        LAUNCH(h0, k_axpy, L0.gn, 256, L0.g, L0.sol, (const double*)L0.cor, 1.0);
        // interpolationAmr: cor1 = trilinear interpolant of cor0 on EVERY fine node; sol1 += cor1
        CK(cudaMemsetAsync(L1.cor, 0, bytes1, h0->stream));
        if (h1->interp_tz == 4)
            launch_pdl(h0, k_interp_tile<false, 4>, L1.git, dim3(256), it_v_doubles(4) * sizeof(double), gN, C->view, L1.cor,
                       (const double*)(L0.cor + C->off_n));
        else
            launch_pdl(h0, k_interp_tile<false>, L1.git, dim3(256), IT_V_DOUBLES * sizeof(double), gN, C->view, L1.cor,
                       (const double*)(L0.cor + C->off_n));
        LAUNCH(h0, k_axpy, L1.gn, 256, gD, L1.sol, (const double*)L1.cor, 1.0);
        residual(h1, L1, L1.sol, L1.rhs, L1.res, nullptr);
        CK(cudaMemsetAsync(L1.cor, 0, bytes1, h0->stream));
        x = L1.cor; y = L1.cor2;
        smooth_sweeps(h1, L1, x, y, L1.res, h0->opts.num_post_smooth * nsw, true);
        LAUNCH(h0, k_axpy, L1.gn, 256, gD, L1.sol, (const double*)x, 1.0);
        // convergence on the composite residual
        residual(h1, L1, L1.sol, L1.rhs, L1.res, h1->partial);
        comp_coarse_residual(C, L0.sol, L1.sol, sums);
        st->resnorm = comp_read_norm(C, nb1);
        st->iters = it + 1;
        if (it + 1 < 128) st->resnorm_hist[it + 1] = st->resnorm;
        if (h0->opts.verbose >= 2) printf("MLMG: Iteration %3d Fine resid/bnorm = %.12g\n", it + 1, st->resnorm / maxnorm);
        if (st->resnorm <= target) { converged = true; break; }
        if (!(st->resnorm <= 1e20 * maxnorm)) { st->status = B200NP_ERR_DIVERGED; break; }
    }
    if (!converged && st->status == B200NP_OK) st->status = B200NP_ERR_NOT_CONVERGED;
    CK(cudaMemcpyAsync(h0->hinfo, h0->dinfo, 4 * sizeof(int), cudaMemcpyDeviceToHost, h0->stream));
    CK(cudaStreamSynchronize(h0->stream));
    st->bottom_iters = h0->hinfo[0];
    if (talk) printf("MLMG: Final Iter. %d resid, resid/bnorm = %.12g, %.12g\n", st->iters, st->resnorm, st->resnorm / maxnorm);
    CK(cudaEventRecord(h0->ev[3], h0->stream));
    // ---- finish: injection, u -= sigma G phi, gphi = G phi, average_down onto the covered cells ----
    LAUNCH(h0, k_comp_inject, box_grid(b.nbn[0], b.nbn[1], b.nbn[2]), 256, L0.g, gD, b, L0.sol, (const double*)L1.sol);
    // NodalProjector::project averages the projected velocity down; ApplyNodalProjection (:84-91) adds velocity_o back
    // afterwards, level by level -- so covered coarse cells end up with avg(u1) + uo0, not avg(u1 + uo1)
    LAUNCH(h0, k_mknewu, L1.gc, 256, gD, (const double*)L1.sol, vel1, Fab{}, 0, gphi1, acc_g);
    LAUNCH(h0, k_mknewu, L0.gc, 256, L0.g, (const double*)L0.sol, vel0, Fab{}, 0, gphi0, acc_g);
    LAUNCH(h0, k_comp_avgdown, gbox, 256, b, vel1, vel0, 3);
    if (add_old) {
        LAUNCH(h0, k_add_cells, L1.gc, 256, gD, vel1, velo1, 3);
        LAUNCH(h0, k_add_cells, L0.gc, 256, L0.g, vel0, velo0, 3);
    }
    if (gphi0.p && gphi1.p) LAUNCH(h0, k_comp_avgdown, gbox, 256, b, gphi1, gphi0, 3);
    if (p1.p) LAUNCH(h0, k_copy_phi, dim3((p1.nx + 63) / 64, (p1.ny + 3) / 4, p1.nz), 256, gN, (const double*)L1.sol, p1, acc_p);
    if (p0.p) LAUNCH(h0, k_copy_phi, dim3((p0.nx + 63) / 64, (p0.ny + 3) / 4, p0.nz), 256, L0.g, (const double*)L0.sol, p0, acc_p);
    return st->status;
}

// caller's fine-level box (fine index space) -> Fab indexed relative to the fine box
Fab make_fab_fine(const b200np_composite* C, double* p, const b200np_fab* bx)
{
    Fab f = make_fab(p, bx);
    if (bx) for (int d = 0; d < 3; ++d) f.lo[d] -= 2 * C->box.lo[d];
    return f;
}
bool fine_box_ok(const b200np_composite* C, const b200np_fab* bx, int ncomp, int grow, bool nodal)
{
    if (!bx || bx->ncomp < ncomp) return false;
    for (int d = 0; d < 3; ++d) {
        const int lo = 2 * C->box.lo[d], hi = 2 * C->box.hi[d] + 1 + (nodal ? 1 : 0);
        const int gr = C->box.span[d] ? 0 : grow;   // periodic ghosts are never read (the library wraps)
        if (bx->lo[d] > lo - gr || bx->hi[d] < hi + gr) return false;
    }
    return true;
}

}  // namespace

// ============================================================================================
// C ABI
// ============================================================================================
extern "C" {

void b200np_default_opts(b200np_opts* o)
{
    memset(o, 0, sizeof(*o));
    o->verbose = 0; o->bottom_verbose = 0;
    o->maxiter = 100; o->bottom_maxiter = 100;
    o->bottom_rtol = 1e-4; o->bottom_atol = -1.0;
    o->mg_max_coarsening_level = 100;  // src/incflo.H:458
    o->num_pre_smooth = 2; o->num_post_smooth = 2; o->smooth_num_sweeps = 4;
    o->bottom_solver = 0;
    o->tile[0] = NP_TX; o->tile[1] = NP_TY; o->tile[2] = 64;
    o->use_graph = 1;
}

int b200np_version(void) { return B200NP_VERSION; }

const char* b200np_strerror(int s)
{
    switch (s) {
    case B200NP_OK: return "ok";
    case B200NP_ERR_NOT_CONVERGED: return "MLMG failed to converge within maxiter";
    case B200NP_ERR_DIVERGED: return "MLMG is diverging";
    case B200NP_ERR_BAD_BC: return "get_projection_bc: undefined BC type";
    case B200NP_ERR_BAD_ARG: return "bad argument";
    case B200NP_ERR_CUDA: return "CUDA error / no usable sm_100 device (there is no CPU fallback)";
    case B200NP_ERR_NCCL: return "NCCL error";
    case B200NP_ERR_UNSUPPORTED: return "not supported by this build";
    case B200NP_ERR_PEER_TIMEOUT: return "slab halo exchange timed out: a neighbour rank never raised its flag";
    case B200NP_ERR_INOUT_FLUX: return "cannot enforce solvability: inflow without outflow through the direction_dependent faces, or the reverse";
    default: return "unknown status";
    }
}

int b200np_create(b200np_t** out, const b200np_geom* geom, const b200np_opts* opts, int device)
{
    return create_common(out, geom, opts, device, 0, 1, nullptr);
}

int b200np_create_dist(b200np_t** out, const b200np_geom* geom, const b200np_opts* opts, int device, int rank,
                       int nranks, const void* nccl_unique_id)
{
    return create_common(out, geom, opts, device, rank, nranks, nccl_unique_id);
}

int b200np_nccl_unique_id(void* out128)
{
    if (!out128) return B200NP_ERR_BAD_ARG;
    if (!g_nccl.load()) return B200NP_ERR_NCCL;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return B200NP_ERR_NCCL;
    memcpy(out128, &id, sizeof(id));
    return B200NP_OK;
}

int b200np_slab_range(const b200np_geom* geom, int rank, int nranks, int* cell_lo, int* cell_hi, int* node_lo, int* node_hi)
{
    if (!geom || nranks < 1 || rank < 0 || rank >= nranks || geom->n_cell[2] % nranks != 0) return B200NP_ERR_BAD_ARG;
    const int m = geom->n_cell[2] / nranks;
    const bool per = geom->bc_lo[2] == B200NP_BC_PERIODIC;
    if (cell_lo) *cell_lo = rank * m;
    if (cell_hi) *cell_hi = (rank + 1) * m - 1;
    if (node_lo) *node_lo = rank * m;
    if (node_hi) *node_hi = (rank + 1) * m - 1 + ((!per && rank == nranks - 1) ? 1 : 0);  // owned (unique) node planes
    return B200NP_OK;
}

int b200np_dist_plan(const b200np_geom* geom, int nranks, int min_planes, int max_coarsening_level, int* nlev_dist, int* nlev)
{
    if (!geom || nranks < 1) return B200NP_ERR_BAD_ARG;
    int rc = check_geom(geom);
    if (rc) return rc;
    if (min_planes <= 0) {
        min_planes = 64;
        if (const char* e = getenv("B200NP_DIST_MIN_PLANES")) min_planes = std::max(8, atoi(e));
    }
    int n[3] = {geom->n_cell[0], geom->n_cell[1], geom->n_cell[2]};
    int lev = 0, nd = 0;
    bool still = nranks > 1;
    for (;;) {
        if (still && level_stays_dist(n, nranks, lev, min_planes)) nd = lev + 1;
        else {
            if (still && lev == 0) return B200NP_ERR_BAD_ARG;   // level 0 must be distributable
            still = false;
        }
        ++lev;
        if (!can_coarsen(n, lev, max_coarsening_level)) break;
        for (int d = 0; d < 3; ++d) n[d] /= 2;
    }
    if (nranks > 1 && nd >= lev) return B200NP_ERR_BAD_ARG;   // needs a replicated coarse level
    if (nlev_dist) *nlev_dist = nd;
    if (nlev) *nlev = lev;
    return B200NP_OK;
}

void b200np_destroy(b200np_t* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->p2p) {  // the neighbours may still be pulling from this arena: handshake before unmapping / freeing
        try { p2p_fence(h); } catch (int) {}
        cudaStreamSynchronize(h->stream);
        if (h->peer_map == 1) {
            b200np_peer::vmm_free(g_drv, h->peer_lo_vmm);
            b200np_peer::vmm_free(g_drv, h->peer_hi_vmm);
        } else {
            if (h->peer_lo) cudaIpcCloseMemHandle(h->peer_lo);
            if (h->peer_hi && h->peer_hi != h->peer_lo) cudaIpcCloseMemHandle(h->peer_hi);
        }
        h->p2p = false;
    }
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
    if (h->graph) cudaGraphDestroy(h->graph);
    if (h->comm) g_nccl.CommDestroy(h->comm);
    if (h->arena_vmm.mapped) b200np_peer::vmm_free(g_drv, h->arena_vmm);
    else if (h->arena.base) cudaFree(h->arena.base);
    if (h->hscal) cudaFreeHost(h->hscal);
    if (h->hinfo) cudaFreeHost(h->hinfo);
    for (auto& s : h->stage) if (s.d) cudaFree(s.d);
    for (auto& m : h->mf) { if (m.slab) cudaFree(m.slab); if (m.tab) cudaFree(m.tab); if (m.stage) cudaFree(m.stage); }
    for (auto& e : h->mf_ev) if (e) cudaEventDestroy(e);
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    for (auto& e : h->prof_ev) cudaEventDestroy(e);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

int b200np_set_stream(b200np_t* h, void* stream)
{
    if (!h) return B200NP_ERR_BAD_ARG;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    h->stream = stream ? (cudaStream_t)stream : h->own_stream;
    return B200NP_OK;
}

// device-side face kinds of the profile: INFLOW faces are mass inflow unless b200np_set_face_types said otherwise
static void sync_profile_faces(b200np* h)
{
    for (int o = 0; o < 6; ++o) {
        const int d = o % 3, bc = o < 3 ? h->geom.bc_lo[d] : h->geom.bc_hi[d];
        int ft = FACE_PLAIN;
        if (bc == B200NP_BC_INFLOW)
            ft = h->face_type[o] == B200NP_FACE_DIRECTION_DEPENDENT ? FACE_DIRECTION_DEPENDENT
               : h->face_type[o] == B200NP_FACE_MIXED ? FACE_MIXED : FACE_MASS_INFLOW;
        h->profile_data.face[o] = ft;
    }
}

int b200np_set_face_types(b200np_t* h, const int face_type[6], int mixed_split_dir, int mixed_half_num_cells)
{
    if (!h || !face_type) return B200NP_ERR_BAD_ARG;
    int mixm = 0;
    bool dd = false;
    for (int o = 0; o < 6; ++o) {
        const int d = o % 3, bc = o < 3 ? h->geom.bc_lo[d] : h->geom.bc_hi[d];
        if (face_type[o] < B200NP_FACE_DEFAULT || face_type[o] > B200NP_FACE_MIXED) return B200NP_ERR_BAD_ARG;
        // get_projection_bc maps both kinds to LinOpBCType::inflow (incflo_projection_bc.cpp:21-27)
        if (face_type[o] != B200NP_FACE_DEFAULT && bc != B200NP_BC_INFLOW) return B200NP_ERR_BAD_BC;
        if (face_type[o] == B200NP_FACE_MIXED) mixm |= 1 << o;
        if (face_type[o] == B200NP_FACE_DIRECTION_DEPENDENT) dd = true;
    }
    if (mixm) {
        if (mixed_split_dir < 0 || mixed_split_dir > 2) return B200NP_ERR_BAD_ARG;
        if (mixed_half_num_cells < 0 || mixed_half_num_cells > h->geom.n_cell[mixed_split_dir]) return B200NP_ERR_BAD_ARG;
        for (int o = 0; o < 6; ++o)   // prob_set_BC_MF splits a face along a direction inside the face
            if ((mixm >> o & 1) && o % 3 == mixed_split_dir) return B200NP_ERR_BAD_ARG;
    }
    for (int o = 0; o < 6; ++o) h->face_type[o] = face_type[o];
    h->has_dd = dd;
    // the overset mask of the mixed faces on every multigrid level (injection: half >> level)
    for (size_t l = 0; l < h->lv.size(); ++l) {
        for (Lev* g : {&h->lv[l].g, &h->lv[l].gpart}) {
            g->mixm = mixm; g->mixdir = mixm ? mixed_split_dir : 0; g->mixhalf = mixm ? mixed_half_num_cells >> l : 0;
        }
    }
    h->singular = 1;
    for (int d = 0; d < 3; ++d)
        if (h->geom.bc_lo[d] == B200NP_BC_DIRICHLET || h->geom.bc_hi[d] == B200NP_BC_DIRICHLET) h->singular = 0;
    if (mixm) h->singular = 0;
    // the level descriptors are kernel parameters baked into the captured V-cycle
    if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); cudaGraphDestroy(h->graph); h->graph_exec = nullptr; h->graph = nullptr; }
    sync_profile_faces(h);
    return B200NP_OK;
}

int b200np_check_overset_mask(b200np_t* h, const int* mask, const b200np_fab* mask_box)
{
    if (!h || !mask || !mask_box) return B200NP_ERR_BAD_ARG;
    try {
        CK(cudaSetDevice(h->device));
        const int nx = mask_box->hi[0] - mask_box->lo[0] + 1, ny = mask_box->hi[1] - mask_box->lo[1] + 1, nz = mask_box->hi[2] - mask_box->lo[2] + 1;
        if (nx < 1 || ny < 1 || nz < 1) return B200NP_ERR_BAD_ARG;
        const size_t bytes = (size_t)nx * ny * nz * sizeof(int);
        const int* d = mask;
        int* staged = nullptr;
        if (!is_device_ptr(mask)) {
            CK(cudaMalloc(&staged, bytes));
            CK(cudaMemcpyAsync(staged, mask, bytes, cudaMemcpyHostToDevice, h->stream));
            d = staged;
        }
        Lev Lm = h->lv[0].g;
        for (int dd = 0; dd < 3; ++dd) Lm.dlo[dd] = Lm.dhi[dd] = 0;
        CK(cudaMemsetAsync(h->dinfo + 8, 0, sizeof(int), h->stream));
        const long long total = (long long)nx * ny * nz;
        LAUNCH(h, k_check_overset, (int)std::min<long long>((total + 255) / 256, 148 * 8), 256, Lm, d, mask_box->lo[0], mask_box->lo[1],
               mask_box->lo[2], nx, ny, nz, h->dinfo + 8);
        CK(cudaMemcpyAsync(h->hinfo + 8, h->dinfo + 8, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (staged) CK(cudaFree(staged));
        return h->hinfo[8] == 0 ? B200NP_OK : B200NP_ERR_UNSUPPORTED;
    } catch (int e) { return e; }
}

int b200np_inout_flux(const b200np_t* h, double* influx, double* outflux)
{
    if (!h) return B200NP_ERR_BAD_ARG;
    if (influx) *influx = h->inout_flux[0];
    if (outflux) *outflux = h->inout_flux[1];
    return B200NP_OK;
}

int b200np_set_inflow_profile(b200np_t* h, int probtype, const double* bcv_vel, double time)
{
    if (!h) return B200NP_ERR_BAD_ARG;
    if (!bcv_vel) { h->has_profile = false; return B200NP_OK; }
    sync_profile_faces(h);
    h->profile_data.probtype = probtype;
    h->profile_data.time = time;
    for (int o = 0; o < 6; ++o)
        for (int c = 0; c < 3; ++c) h->profile_data.bcv[o][c] = bcv_vel[3 * o + c];
    h->has_profile = true;
    return B200NP_OK;
}

int b200np_nlevels(const b200np_t* h) { return h ? (int)h->lv.size() : 0; }

int b200np_halo_transport(const b200np_t* h) { return !h ? -1 : h->nranks == 1 ? 0 : h->p2p ? 1 : 2; }
int b200np_peer_map(const b200np_t* h) { return !h ? -1 : h->p2p ? h->peer_map : 0; }

int b200np_level_dims(const b200np_t* h, int lev, int n_cell[3], int n_node[3])
{
    if (!h || lev < 0 || lev >= (int)h->lv.size()) return B200NP_ERR_BAD_ARG;
    for (int d = 0; d < 3; ++d) { n_cell[d] = h->lv[lev].g.n[d]; n_node[d] = h->lv[lev].g.nn[d]; }
    n_cell[2] = h->lv[lev].g.cnzl; n_node[2] = h->lv[lev].g.nzl;  // locally owned planes
    return B200NP_OK;
}

int b200np_set_sigma(b200np_t* h, const double* sigma, const b200np_fab* sigma_box, double const_sigma)
{
    if (!h) return B200NP_ERR_BAD_ARG;
    try {
        CK(cudaSetDevice(h->device));
        LevelData& L0 = h->lv[0];
        if (sigma) {
            const int lo[3] = {0, 0, L0.g.ck0}, hi[3] = {L0.g.n[0] - 1, L0.g.n[1] - 1, L0.g.ck0 + L0.g.cnzl - 1};
            if (!box_covers(sigma_box, lo, hi, 1)) return B200NP_ERR_BAD_ARG;
            b200np_stats st{};
            bool staged;
            double* d = stage_in(h, 3, sigma, sigma_box, true, &staged, &st);
            set_level_sigma_ptrs(h, true, 1.0);
            LAUNCH(h, k_copy_sigma, L0.gc, 256, L0.g, make_fab(d, sigma_box), L0.sigma);
        } else {
            set_level_sigma_ptrs(h, false, const_sigma);
        }
        coarsen_sigma(h);
        CK(cudaStreamSynchronize(h->stream));
    } catch (int e) { return e; }
    return B200NP_OK;
}

int b200np_project(b200np_t* h, double* vel, const b200np_fab* vel_box, const double* sigma, const b200np_fab* sigma_box,
                   double const_sigma, double* phi, const b200np_fab* phi_box, double* gphi, const b200np_fab* gphi_box,
                   double rtol, double atol, b200np_stats* stats)
{
    b200np_stats local{};
    b200np_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof(*st));
    if (!h || !vel || !vel_box) return st->status = B200NP_ERR_BAD_ARG;
    try {
        CK(cudaSetDevice(h->device));
        LevelData& L0 = h->lv[0];
        const Lev& g = L0.g;
        const int clo[3] = {0, 0, g.ck0}, chi[3] = {g.n[0] - 1, g.n[1] - 1, g.ck0 + g.cnzl - 1};
        if (!vel_box_ok(h, vel_box)) return st->status = B200NP_ERR_BAD_ARG;
        if (sigma && !box_covers(sigma_box, clo, chi, 1)) return st->status = B200NP_ERR_BAD_ARG;
        if (gphi && !box_covers(gphi_box, clo, chi, 3)) return st->status = B200NP_ERR_BAD_ARG;
        if (phi && !nodal_box_ok(h, phi_box)) return st->status = B200NP_ERR_BAD_ARG;
        h->launches = 0; h->exchanges = 0;
        CK(cudaEventRecord(h->ev[0], h->stream));
        bool s_vel, s_sig, s_phi = false, s_g = false;
        double* dvel = stage_in(h, 0, vel, vel_box, true, &s_vel, st);
        const double* dsig = stage_in(h, 3, sigma, sigma_box, true, &s_sig, st);
        double* dphi = phi ? stage_in(h, 5, phi, phi_box, false, &s_phi, st) : nullptr;
        double* dg = gphi ? stage_in(h, 4, gphi, gphi_box, false, &s_g, st) : nullptr;
        CK(cudaEventRecord(h->ev[4], h->stream));
        if (sigma) {
            set_level_sigma_ptrs(h, true, 1.0);
            LAUNCH(h, k_copy_sigma, L0.gc, 256, L0.g, make_fab(const_cast<double*>(dsig), sigma_box), L0.sigma);
        } else {
            set_level_sigma_ptrs(h, false, const_sigma);
        }
        int status = project_core(h, make_fab(dvel, vel_box), Fab{}, 0, make_fab(dg, gphi_box), 0,
                                  make_fab(dphi, phi_box), 0, rtol, atol, st);
        CK(cudaEventRecord(h->ev[5], h->stream));
        stage_out(h, 0, vel, vel_box, s_vel, st);
        stage_out(h, 5, phi, phi_box, s_phi, st);
        stage_out(h, 4, gphi, gphi_box, s_g, st);
        finish_stats(h, st);
        float ms;
        CK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[4])); st->ms_h2d = ms;
        CK(cudaEventElapsedTime(&ms, h->ev[5], h->ev[1])); st->ms_d2h = ms;
        return st->status = status;
    } catch (int e) { return st->status = e; }
}

// HydroUtils::enforceInOutSolvability (call site incflo_apply_nodal_projection.cpp:166-179) on the ghost layer of
// the direction_dependent faces.  AMReX-Hydro aborts when only one of influx / outflux is non-zero; here that is
// B200NP_ERR_INOUT_FLUX.
namespace {
int enforce_inout_solvability(b200np* h, Fab fvel)
{
    constexpr double small_vel = 1.0e-8;
    const Lev& g = h->lv[0].g;
    sync_profile_faces(h);
    const int blocks = 148;
    LAUNCH(h, k_inout_flux, blocks, 256, g, fvel, h->profile_data, h->partial);
    LAUNCH(h, k_sum2_final, 1, 1024, h->partial, (long long)blocks, h->dscal);
    allreduce(h, h->dscal, 2, ncclSum);
    CK(cudaMemcpyAsync(h->hscal, h->dscal, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    LAUNCH(h, k_inout_correct, blocks, 256, g, fvel, h->profile_data, h->dscal, small_vel);
    CK(cudaStreamSynchronize(h->stream));
    h->inout_flux[0] = h->hscal[0]; h->inout_flux[1] = h->hscal[1];
    const bool in = h->hscal[0] > small_vel, out = h->hscal[1] > small_vel;
    if (in != out) return B200NP_ERR_INOUT_FLUX;   // "Cannot enforce solvability": inflow without outflow, or the reverse
    return B200NP_OK;
}
}  // namespace

int b200np_apply_nodal_projection(b200np_t* h, double* velocity, const b200np_fab* vel_box, const double* velocity_o,
                                  const double* density, const b200np_fab* rho_box, double ro_0, double* gp,
                                  const b200np_fab* gp_box, double* p_nd, const b200np_fab* p_box, const double* inflow_vel,
                                  double scaling_factor, int incremental, int proj_for_small_dt, double rtol, double atol,
                                  b200np_stats* stats)
{
    b200np_stats local{};
    b200np_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof(*st));
    if (!h || !velocity || !vel_box || !gp || !gp_box || !p_nd || !p_box) return st->status = B200NP_ERR_BAD_ARG;
    const int use_old = (incremental || proj_for_small_dt);
    if (use_old && !velocity_o) return st->status = B200NP_ERR_BAD_ARG;
    try {
        CK(cudaSetDevice(h->device));
        LevelData& L0 = h->lv[0];
        const Lev& g = L0.g;
        const int clo[3] = {0, 0, g.ck0}, chi[3] = {g.n[0] - 1, g.n[1] - 1, g.ck0 + g.cnzl - 1};
        if (!vel_box_ok(h, vel_box) || !box_covers(gp_box, clo, chi, 3)) return st->status = B200NP_ERR_BAD_ARG;
        if (density && !box_covers(rho_box, clo, chi, 1)) return st->status = B200NP_ERR_BAD_ARG;
        if (!nodal_box_ok(h, p_box)) return st->status = B200NP_ERR_BAD_ARG;
        h->launches = 0; h->exchanges = 0;
        CK(cudaEventRecord(h->ev[0], h->stream));
        bool s_vel, s_velo, s_rho, s_gp, s_p, s_in;
        double* dvel = stage_in(h, 0, velocity, vel_box, true, &s_vel, st);
        double* dvelo = stage_in(h, 1, use_old ? velocity_o : nullptr, vel_box, true, &s_velo, st);
        double* drho = stage_in(h, 2, density, rho_box, true, &s_rho, st);
        double* dgp = stage_in(h, 4, gp, gp_box, true, &s_gp, st);   // gp is an input in both modes
        double* dp = stage_in(h, 5, p_nd, p_box, incremental != 0, &s_p, st);
        const int set_inflow = (!proj_for_small_dt && !incremental);   // :81
        double* din = stage_in(h, 6, set_inflow ? inflow_vel : nullptr, vel_box, true, &s_in, st);
        CK(cudaEventRecord(h->ev[4], h->stream));
        Fab fvel = make_fab(dvel, vel_box), fvelo = make_fab(dvelo, vel_box), frho = make_fab(drho, rho_box),
            fgp = make_fab(dgp, gp_box), fp = make_fab(dp, p_box), fin = make_fab(din, vel_box);
        // :39-71, :101-121 fused
        set_level_sigma_ptrs(h, density != nullptr, scaling_factor / ro_0);
        if (!incremental || use_old || density)
            LAUNCH(h, k_pre_add_sigma, L0.gc, 256, L0.g, fvel, fgp, frho, fvelo, scaling_factor, ro_0, incremental ? 0 : 1,
                   use_old, density ? L0.sigma : nullptr);
        // :137-163
        {
            long long total = (long long)fvel.nx * fvel.ny * fvel.nz;
            int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
            LAUNCH(h, k_set_vel_ghosts, blocks, 256, L0.g, fvel, fin, set_inflow);
            // no caller-filled array: IncfloVelFill on the device (:138-163), when a profile has been set
            if (set_inflow && !fin.p && h->has_profile) LAUNCH(h, k_incflo_vel_fill, blocks, 256, L0.g, fvel, h->profile_data);
            // :166-179 enforceInOutSolvability over the direction_dependent faces
            if (set_inflow && h->has_dd) {
                int rc = enforce_inout_solvability(h, fvel);
                if (rc) return st->status = rc;
            }
        }
        // :181-256
        int status = project_core(h, fvel, fvelo, use_old, fgp, incremental, fp, incremental, rtol, atol, st);
        CK(cudaEventRecord(h->ev[5], h->stream));
        stage_out(h, 0, velocity, vel_box, s_vel, st);
        stage_out(h, 4, gp, gp_box, s_gp, st);
        stage_out(h, 5, p_nd, p_box, s_p, st);
        finish_stats(h, st);
        float ms;
        CK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[4])); st->ms_h2d = ms;
        CK(cudaEventElapsedTime(&ms, h->ev[5], h->ev[1])); st->ms_d2h = ms;
        return st->status = status;
    } catch (int e) { return st->status = e; }
}


// ---- multi-box MultiFabs (b200np_*_mf): gather into the slab arrays, run the single-box path, scatter back ----
namespace {
struct MfField {
    const b200np_mfab* mf = nullptr;
    MfFab* tab = nullptr;            // device table of the fabs (device pointers)
    std::vector<MfFab> host;
    std::vector<size_t> off;         // staged: offset of fab f in the staging area (doubles)
    bool staged = false;
};
size_t mf_fab_doubles(const b200np_fab& b, int ncomp)
{
    return (size_t)(b.hi[0] - b.lo[0] + 1) * (b.hi[1] - b.lo[1] + 1) * (b.hi[2] - b.lo[2] + 1) * ncomp;
}
// the valid boxes must lie inside [lo, hi] (cell or nodal index space of the rank) and, for cell-centred MultiFabs,
// tile it (the volumes add up; amrex::BoxArray boxes never overlap)
bool mf_boxes_ok(const b200np_mfab* m, const int lo[3], const int hi[3], int ncomp, bool tile)
{
    if (!m || m->nfabs < 1 || m->ngrow < 0 || m->ncomp < ncomp || !m->box || !m->data) return false;
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
void mf_map(b200np* h, int slot, const b200np_mfab* m, int ncomp, bool copy_in, b200np_stats* st, MfField& F)
{
    F = MfField{};
    if (!m) return;
    F.mf = m;
    auto& S = h->mf[slot];
    const int nf = m->nfabs;
    F.staged = !is_device_ptr(m->data[0]);
    F.host.resize(nf); F.off.assign(nf, 0);
    size_t total = 0;
    for (int f = 0; f < nf; ++f) { F.off[f] = total; total += mf_fab_doubles(m->box[f], m->ncomp); }
    if (F.staged) {
        if (S.stage_bytes < total * sizeof(double)) {
            if (S.stage) CK(cudaFree(S.stage));
            CK(cudaMalloc(&S.stage, total * sizeof(double)));
            S.stage_bytes = total * sizeof(double);
        }
        if (copy_in) {
            for (int f = 0; f < nf; ++f)
                CK(cudaMemcpyAsync(S.stage + F.off[f], m->data[f], mf_fab_doubles(m->box[f], m->ncomp) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            st->h2d_bytes += (long long)(total * sizeof(double));
        }
    }
    for (int f = 0; f < nf; ++f) {
        MfFab& d = F.host[f];
        const b200np_fab& b = m->box[f];
        d.p = F.staged ? S.stage + F.off[f] : m->data[f];
        for (int q = 0; q < 3; ++q) d.lo[q] = b.lo[q];
        d.nx = b.hi[0] - b.lo[0] + 1; d.ny = b.hi[1] - b.lo[1] + 1; d.nz = b.hi[2] - b.lo[2] + 1;
        d.cstride = (long long)d.nx * d.ny * d.nz;
    }
    if (S.tab_cap < (size_t)nf) {
        if (S.tab) CK(cudaFree(S.tab));
        CK(cudaMalloc(&S.tab, (size_t)nf * sizeof(MfFab)));
        S.tab_cap = nf;
    }
    CK(cudaMemcpyAsync(S.tab, F.host.data(), (size_t)nf * sizeof(MfFab), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));   // F.host is about to go out of the caller's sight; tables are tiny
    F.tab = S.tab;
    (void)ncomp;
}
void mf_copy_back(b200np* h, int slot, MfField& F, b200np_stats* st)
{
    if (!F.mf || !F.staged) return;
    auto& S = h->mf[slot];
    size_t total = 0;
    for (int f = 0; f < F.mf->nfabs; ++f) {
        const size_t nd = mf_fab_doubles(F.mf->box[f], F.mf->ncomp);
        CK(cudaMemcpyAsync(F.mf->data[f], S.stage + F.off[f], nd * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        total += nd;
    }
    st->d2h_bytes += (long long)(total * sizeof(double));
}
double* mf_slab(b200np* h, int slot, const b200np_fab& box)
{
    auto& S = h->mf[slot];
    const size_t bytes = fab_bytes(&box);
    if (S.slab_bytes < bytes) {
        if (S.slab) CK(cudaFree(S.slab));
        CK(cudaMalloc(&S.slab, bytes));
        S.slab_bytes = bytes;
    }
    return S.slab;
}
dim3 mf_grid(const MfField& F)
{
    long long mx = 1;
    for (const MfFab& f : F.host) mx = std::max(mx, (long long)f.nx * f.ny * f.nz);
    return dim3((unsigned)std::min<long long>((mx + 255) / 256, 64), (unsigned)F.host.size());
}
void mf_gather(b200np* h, const MfField& F, int ncomp, double* slab, const b200np_fab& box, int mode)
{
    if (!F.mf) return;
    const Lev& g = h->lv[0].g;
    LAUNCH(h, k_mf_gather, mf_grid(F), 256, F.tab, F.mf->ngrow, ncomp, make_fab(slab, &box), mode, g.n[0], g.n[1], g.n[2]);
}
void mf_scatter(b200np* h, const MfField& F, int ncomp, double* slab, const b200np_fab& box, int mode)
{
    if (!F.mf) return;
    LAUNCH(h, k_mf_scatter, mf_grid(F), 256, F.tab, F.mf->ngrow, ncomp, make_fab(slab, &box), mode);
}
// the slab boxes of this rank: cells, cells grown by one (velocity), nodes
void mf_slab_boxes(const b200np* h, b200np_fab& cells, b200np_fab& grown, b200np_fab& nodes)
{
    const Lev& g = h->lv[0].g;
    const int lo[3] = {0, 0, g.ck0}, hi[3] = {g.n[0] - 1, g.n[1] - 1, g.ck0 + g.cnzl - 1};
    for (int d = 0; d < 3; ++d) {
        cells.lo[d] = lo[d]; cells.hi[d] = hi[d];
        grown.lo[d] = lo[d] - 1; grown.hi[d] = hi[d] + 1;
        nodes.lo[d] = lo[d]; nodes.hi[d] = hi[d] + 1;
    }
    cells.ncomp = 1; grown.ncomp = 3; nodes.ncomp = 1;
}
void mf_events(b200np* h)
{
    for (auto& e : h->mf_ev) if (!e) CK(cudaEventCreate(&e));
}
}  // namespace

int b200np_project_mf(b200np_t* h, const b200np_mfab* vel, const b200np_mfab* sigma, double const_sigma, const b200np_mfab* phi,
                      const b200np_mfab* gphi, double rtol, double atol, b200np_stats* stats)
{
    b200np_stats local{};
    b200np_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof(*st));
    if (!h) return st->status = B200NP_ERR_BAD_ARG;
    try {
        CK(cudaSetDevice(h->device));
        b200np_fab cells{}, grown{}, nodes{};
        mf_slab_boxes(h, cells, grown, nodes);
        if (!mf_boxes_ok(vel, cells.lo, cells.hi, 3, true) || vel->ngrow < 1) return st->status = B200NP_ERR_BAD_ARG;
        if (sigma && !mf_boxes_ok(sigma, cells.lo, cells.hi, 1, true)) return st->status = B200NP_ERR_BAD_ARG;
        if (gphi && !mf_boxes_ok(gphi, cells.lo, cells.hi, 3, true)) return st->status = B200NP_ERR_BAD_ARG;
        if (phi && !mf_boxes_ok(phi, nodes.lo, nodes.hi, 1, false)) return st->status = B200NP_ERR_BAD_ARG;
        mf_events(h);
        b200np_stats outer{};
        CK(cudaEventRecord(h->mf_ev[0], h->stream));
        MfField Fv, Fs, Fp, Fg;
        mf_map(h, 0, vel, 3, true, &outer, Fv);
        mf_map(h, 2, sigma, 1, true, &outer, Fs);
        mf_map(h, 4, phi, 1, false, &outer, Fp);
        mf_map(h, 3, gphi, 3, false, &outer, Fg);
        b200np_fab gbox = cells; gbox.ncomp = 3;
        double* svel = mf_slab(h, 0, grown);
        double* ssig = sigma ? mf_slab(h, 2, cells) : nullptr;
        double* sphi = phi ? mf_slab(h, 4, nodes) : nullptr;
        double* sg = gphi ? mf_slab(h, 3, gbox) : nullptr;
        long long extra = h->launches;
        h->launches = 0;
        CK(cudaMemsetAsync(svel, 0, fab_bytes(&grown), h->stream));
        mf_gather(h, Fv, 3, svel, grown, MF_VALID_BC);
        mf_gather(h, Fs, 1, ssig, cells, MF_VALID);
        extra = h->launches;
        int rc = b200np_project(h, svel, &grown, ssig, sigma ? &cells : nullptr, const_sigma, sphi, phi ? &nodes : nullptr, sg,
                                gphi ? &gbox : nullptr, rtol, atol, st);
        if (rc != B200NP_OK && rc != B200NP_ERR_NOT_CONVERGED && rc != B200NP_ERR_DIVERGED) return rc;
        h->launches = 0;
        mf_scatter(h, Fv, 3, svel, grown, MF_VALID);
        mf_scatter(h, Fp, 1, sphi, nodes, MF_VALID);
        mf_scatter(h, Fg, 3, sg, gbox, MF_VALID);
        extra += h->launches;
        mf_copy_back(h, 0, Fv, &outer); mf_copy_back(h, 4, Fp, &outer); mf_copy_back(h, 3, Fg, &outer);
        CK(cudaEventRecord(h->mf_ev[1], h->stream));
        CK(cudaEventSynchronize(h->mf_ev[1]));
        float ms;
        CK(cudaEventElapsedTime(&ms, h->mf_ev[0], h->mf_ev[1])); st->ms_total = ms;
        st->launches += extra; st->h2d_bytes += outer.h2d_bytes; st->d2h_bytes += outer.d2h_bytes;
        return st->status = rc;
    } catch (int e) { return st->status = e; }
}

int b200np_apply_nodal_projection_mf(b200np_t* h, const b200np_mfab* velocity, const b200np_mfab* velocity_o,
                                     const b200np_mfab* density, double ro_0, const b200np_mfab* gp, const b200np_mfab* p_nd,
                                     const b200np_mfab* inflow_vel, double scaling_factor, int incremental, int proj_for_small_dt,
                                     double rtol, double atol, b200np_stats* stats)
{
    b200np_stats local{};
    b200np_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof(*st));
    if (!h) return st->status = B200NP_ERR_BAD_ARG;
    const int use_old = (incremental || proj_for_small_dt);
    try {
        CK(cudaSetDevice(h->device));
        b200np_fab cells{}, grown{}, nodes{};
        mf_slab_boxes(h, cells, grown, nodes);
        if (!mf_boxes_ok(velocity, cells.lo, cells.hi, 3, true) || velocity->ngrow < 1) return st->status = B200NP_ERR_BAD_ARG;
        if (use_old && !mf_boxes_ok(velocity_o, cells.lo, cells.hi, 3, true)) return st->status = B200NP_ERR_BAD_ARG;
        if (density && !mf_boxes_ok(density, cells.lo, cells.hi, 1, true)) return st->status = B200NP_ERR_BAD_ARG;
        if (!mf_boxes_ok(gp, cells.lo, cells.hi, 3, true) || !mf_boxes_ok(p_nd, nodes.lo, nodes.hi, 1, false)) return st->status = B200NP_ERR_BAD_ARG;
        const int set_inflow = (!proj_for_small_dt && !incremental);
        if (inflow_vel && set_inflow && (!mf_boxes_ok(inflow_vel, cells.lo, cells.hi, 3, true) || inflow_vel->ngrow < 1)) return st->status = B200NP_ERR_BAD_ARG;
        mf_events(h);
        b200np_stats outer{};
        CK(cudaEventRecord(h->mf_ev[0], h->stream));
        MfField Fv, Fo, Fr, Fg, Fp, Fi;
        mf_map(h, 0, velocity, 3, true, &outer, Fv);
        mf_map(h, 1, use_old ? velocity_o : nullptr, 3, true, &outer, Fo);
        mf_map(h, 2, density, 1, true, &outer, Fr);
        mf_map(h, 3, gp, 3, true, &outer, Fg);
        mf_map(h, 4, p_nd, 1, incremental != 0, &outer, Fp);
        mf_map(h, 5, (inflow_vel && set_inflow) ? inflow_vel : nullptr, 3, true, &outer, Fi);
        b200np_fab gbox = cells; gbox.ncomp = 3;
        double* svel = mf_slab(h, 0, grown);
        double* svelo = Fo.mf ? mf_slab(h, 1, grown) : nullptr;
        double* srho = Fr.mf ? mf_slab(h, 2, cells) : nullptr;
        double* sgp = mf_slab(h, 3, gbox);
        double* sp = mf_slab(h, 4, nodes);
        double* sin = Fi.mf ? mf_slab(h, 5, grown) : nullptr;
        h->launches = 0;
        mf_gather(h, Fv, 3, svel, grown, MF_VALID);      // every ghost cell is set by setBndry(0) / the inflow fill below
        mf_gather(h, Fo, 3, svelo, grown, MF_VALID);
        mf_gather(h, Fr, 1, srho, cells, MF_VALID);
        mf_gather(h, Fg, 3, sgp, gbox, MF_VALID);
        if (incremental) mf_gather(h, Fp, 1, sp, nodes, MF_VALID);
        if (Fi.mf) { CK(cudaMemsetAsync(sin, 0, fab_bytes(&grown), h->stream)); mf_gather(h, Fi, 3, sin, grown, MF_VALID_BC); }
        long long extra = h->launches;
        int rc = b200np_apply_nodal_projection(h, svel, &grown, svelo, srho, Fr.mf ? &cells : nullptr, ro_0, sgp, &gbox, sp, &nodes, sin,
                                               scaling_factor, incremental, proj_for_small_dt, rtol, atol, st);
        if (rc != B200NP_OK && rc != B200NP_ERR_NOT_CONVERGED && rc != B200NP_ERR_DIVERGED) return rc;
        h->launches = 0;
        mf_scatter(h, Fv, 3, svel, grown, MF_VALID_BC);
        mf_scatter(h, Fg, 3, sgp, gbox, MF_VALID);
        mf_scatter(h, Fp, 1, sp, nodes, MF_VALID);
        extra += h->launches;
        mf_copy_back(h, 0, Fv, &outer); mf_copy_back(h, 3, Fg, &outer); mf_copy_back(h, 4, Fp, &outer);
        CK(cudaEventRecord(h->mf_ev[1], h->stream));
        CK(cudaEventSynchronize(h->mf_ev[1]));
        float ms;
        CK(cudaEventElapsedTime(&ms, h->mf_ev[0], h->mf_ev[1])); st->ms_total = ms;
        st->launches += extra; st->h2d_bytes += outer.h2d_bytes; st->d2h_bytes += outer.d2h_bytes;
        return st->status = rc;
    } catch (int e) { return st->status = e; }
}

// ---- composite (two AMR level) projection -----------------------------------------------------
int b200np_composite_create(b200np_composite_t** out, const b200np_geom* geom0, const int fine_lo[3], const int fine_hi[3],
                            const b200np_opts* opts, int device)
{
    if (!out || !geom0 || !fine_lo || !fine_hi) return B200NP_ERR_BAD_ARG;
    *out = nullptr;
    int rcg = check_geom(geom0);
    if (rcg) return rcg;
    int span[3], cflo[3], cfhi[3], ncf = 0;
    for (int d = 0; d < 3; ++d) {
        if (fine_lo[d] < 0 || fine_hi[d] < fine_lo[d] || fine_hi[d] > geom0->n_cell[d] - 1) return B200NP_ERR_BAD_ARG;
        const bool at_lo = fine_lo[d] == 0, at_hi = fine_hi[d] == geom0->n_cell[d] - 1;
        span[d] = 0; cflo[d] = !at_lo; cfhi[d] = !at_hi;
        if (geom0->bc_lo[d] == B200NP_BC_PERIODIC) {
            if (at_lo != at_hi) return B200NP_ERR_UNSUPPORTED;   // touches the periodic seam without spanning the direction
            span[d] = at_lo;
        } else {   // a fine box on an inflow face would need the inflow profile on the fine level too
            if (at_lo && geom0->bc_lo[d] == B200NP_BC_INFLOW) return B200NP_ERR_UNSUPPORTED;
            if (at_hi && geom0->bc_hi[d] == B200NP_BC_INFLOW) return B200NP_ERR_UNSUPPORTED;
        }
        ncf += cflo[d] + cfhi[d];
    }
    if (ncf == 0) return B200NP_ERR_UNSUPPORTED;   // the "fine box" is the whole domain
    b200np_composite* C = new b200np_composite();
    int rc = create_common(&C->h0, geom0, opts, device, 0, 1, nullptr);
    if (rc) { delete C; return rc; }
    b200np_geom g1 = *geom0;
    b200np_opts o1 = C->h0->opts;
    o1.mg_max_coarsening_level = 0;   // ref ratio 2: the fine AMR level has ONE multigrid level
    for (int d = 0; d < 3; ++d) {
        C->box.lo[d] = fine_lo[d]; C->box.hi[d] = fine_hi[d];
        C->box.cf_lo[d] = cflo[d]; C->box.cf_hi[d] = cfhi[d]; C->box.span[d] = span[d];
        C->nb[d] = fine_hi[d] - fine_lo[d] + 1;
        C->box.nbn[d] = span[d] ? C->nb[d] : C->nb[d] + 1;
        g1.n_cell[d] = 2 * C->nb[d]; g1.dx[d] = 0.5 * geom0->dx[d];
        // interface: Dirichlet (not relaxed); otherwise the fine level inherits the domain's BC on that side
        g1.bc_lo[d] = cflo[d] ? B200NP_BC_DIRICHLET : geom0->bc_lo[d];
        g1.bc_hi[d] = cfhi[d] ? B200NP_BC_DIRICHLET : geom0->bc_hi[d];
    }
    rc = create_common(&C->h1, &g1, &o1, device, 0, 1, nullptr, true);
    if (rc) { b200np_destroy(C->h0); delete C; return rc; }
    try {
        CK(cudaSetDevice(device));
        C->h1->stream = C->h0->stream;   // one stream for both levels
        const Lev& g0 = C->h0->lv[0].g;
        const Lev& gf = C->h1->lv[0].g;
        Lev v = g0;
        for (int d = 0; d < 3; ++d) {
            v.n[d] = C->nb[d]; v.nn[d] = C->box.nbn[d]; v.per[d] = C->box.span[d];
            v.rlo[d] = v.rhi[d] = 0; v.dlo[d] = v.dhi[d] = 0;
        }
        v.k0 = 0; v.nzl = C->box.nbn[2]; v.ck0 = 0; v.cnzl = C->nb[2]; v.dist = 0; v.sigma = nullptr;
        C->view = v;
        C->off_n = (long long)fine_lo[2] * g0.ps + (long long)fine_lo[1] * g0.px + fine_lo[0];
        CK(cudaMalloc(&C->rhsN, (size_t)gf.ps * gf.nzl * sizeof(double)));
        CK(cudaMalloc(&C->s0z_alloc, (size_t)g0.cps * (g0.cnzl + 2) * sizeof(double)));
        CK(cudaMemset(C->s0z_alloc, 0, (size_t)g0.cps * (g0.cnzl + 2) * sizeof(double)));
        C->s0z = C->s0z_alloc + g0.cps;
    } catch (int e) { b200np_composite_destroy(C); return e; }
    *out = C;
    return B200NP_OK;
}

void b200np_composite_destroy(b200np_composite_t* C)
{
    if (!C) return;
    if (C->h0) { cudaSetDevice(C->h0->device); cudaStreamSynchronize(C->h0->stream); }
    if (C->rhsN) cudaFree(C->rhsN);
    if (C->s0z_alloc) cudaFree(C->s0z_alloc);
    if (C->h1) { C->h1->stream = C->h1->own_stream; b200np_destroy(C->h1); }
    if (C->h0) b200np_destroy(C->h0);
    delete C;
}

int b200np_composite_set_stream(b200np_composite_t* C, void* stream)
{
    if (!C) return B200NP_ERR_BAD_ARG;
    int rc = b200np_set_stream(C->h0, stream);
    C->h1->stream = C->h0->stream;
    return rc;
}

b200np_t* b200np_composite_level(b200np_composite_t* C, int amr_level)
{
    return !C ? nullptr : amr_level == 0 ? C->h0 : amr_level == 1 ? C->h1 : nullptr;
}

int b200np_composite_project(b200np_composite_t* C, double* vel0, const b200np_fab* vel0_box, double* vel1,
                             const b200np_fab* vel1_box, const double* sigma0, const b200np_fab* sigma0_box,
                             const double* sigma1, const b200np_fab* sigma1_box, double const_sigma, double* phi0,
                             const b200np_fab* phi0_box, double* phi1, const b200np_fab* phi1_box, double* gphi0,
                             const b200np_fab* gphi0_box, double* gphi1, const b200np_fab* gphi1_box, double rtol, double atol,
                             b200np_stats* stats)
{
    b200np_stats local{};
    b200np_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof(*st));
    if (!C || !vel0 || !vel0_box || !vel1 || !vel1_box) return st->status = B200NP_ERR_BAD_ARG;
    if ((sigma0 == nullptr) != (sigma1 == nullptr)) return st->status = B200NP_ERR_BAD_ARG;
    b200np *h0 = C->h0, *h1 = C->h1;
    try {
        CK(cudaSetDevice(h0->device));
        LevelData &L0 = h0->lv[0], &L1 = h1->lv[0];
        const Lev& g = L0.g;
        const int clo[3] = {0, 0, 0}, chi[3] = {g.n[0] - 1, g.n[1] - 1, g.n[2] - 1};
        if (!vel_box_ok(h0, vel0_box) || !fine_box_ok(C, vel1_box, 3, 1, false)) return st->status = B200NP_ERR_BAD_ARG;
        if (sigma0 && (!box_covers(sigma0_box, clo, chi, 1) || !fine_box_ok(C, sigma1_box, 1, 0, false))) return st->status = B200NP_ERR_BAD_ARG;
        if (gphi0 && !box_covers(gphi0_box, clo, chi, 3)) return st->status = B200NP_ERR_BAD_ARG;
        if (gphi1 && !fine_box_ok(C, gphi1_box, 3, 0, false)) return st->status = B200NP_ERR_BAD_ARG;
        if ((phi0 && !nodal_box_ok(h0, phi0_box)) || (phi1 && !fine_box_ok(C, phi1_box, 1, 0, true))) return st->status = B200NP_ERR_BAD_ARG;
        h0->launches = h1->launches = 0;
        CK(cudaEventRecord(h0->ev[0], h0->stream));
        bool s_v0, s_v1, s_s0, s_s1, s_p0 = false, s_p1 = false, s_g0 = false, s_g1 = false;
        double* dv0 = stage_in(h0, 0, vel0, vel0_box, true, &s_v0, st);
        double* dv1 = stage_in(h1, 0, vel1, vel1_box, true, &s_v1, st);
        const double* ds0 = stage_in(h0, 3, sigma0, sigma0_box, true, &s_s0, st);
        const double* ds1 = stage_in(h1, 3, sigma1, sigma1_box, true, &s_s1, st);
        double* dp0 = phi0 ? stage_in(h0, 5, phi0, phi0_box, false, &s_p0, st) : nullptr;
        double* dp1 = phi1 ? stage_in(h1, 5, phi1, phi1_box, false, &s_p1, st) : nullptr;
        double* dg0 = gphi0 ? stage_in(h0, 4, gphi0, gphi0_box, false, &s_g0, st) : nullptr;
        double* dg1 = gphi1 ? stage_in(h1, 4, gphi1, gphi1_box, false, &s_g1, st) : nullptr;
        CK(cudaEventRecord(h0->ev[4], h0->stream));
        Fab fv1 = make_fab_fine(C, dv1, vel1_box);
        if (sigma0) {
            LAUNCH(h0, k_copy_sigma, L0.gc, 256, L0.g, make_fab(const_cast<double*>(ds0), sigma0_box), L0.sigma);
            LAUNCH(h0, k_copy_sigma, L1.gc, 256, L1.g, make_fab_fine(C, const_cast<double*>(ds1), sigma1_box), L1.sigma);
        }
        comp_set_sigma(C, sigma0 != nullptr, const_sigma);
        {   // vel.setBndry(0.0) on the fine level (:137); the coarse ghost layer is the caller's input as in b200np_project
            long long total = (long long)fv1.nx * fv1.ny * fv1.nz;
            int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
            LAUNCH(h0, k_set_vel_ghosts, blocks, 256, L1.g, fv1, Fab{}, 0);
        }
        int status = comp_core(C, make_fab(dv0, vel0_box), fv1, Fab{}, Fab{}, 0, make_fab(dg0, gphi0_box), make_fab_fine(C, dg1, gphi1_box), 0,
                               make_fab(dp0, phi0_box), make_fab_fine(C, dp1, phi1_box), 0, rtol, atol, st);
        CK(cudaEventRecord(h0->ev[5], h0->stream));
        stage_out(h0, 0, vel0, vel0_box, s_v0, st); stage_out(h1, 0, vel1, vel1_box, s_v1, st);
        stage_out(h0, 5, phi0, phi0_box, s_p0, st); stage_out(h1, 5, phi1, phi1_box, s_p1, st);
        stage_out(h0, 4, gphi0, gphi0_box, s_g0, st); stage_out(h1, 4, gphi1, gphi1_box, s_g1, st);
        finish_stats(h0, st);
        st->launches = h0->launches + h1->launches;
        float ms;
        CK(cudaEventElapsedTime(&ms, h0->ev[0], h0->ev[4])); st->ms_h2d = ms;
        CK(cudaEventElapsedTime(&ms, h0->ev[5], h0->ev[1])); st->ms_d2h = ms;
        return st->status = status;
    } catch (int e) { return st->status = e; }
}

int b200np_composite_apply_nodal_projection(b200np_composite_t* C, double* const velocity[2], const b200np_fab* const vel_box[2],
                                            const double* const velocity_o[2], const double* const density[2],
                                            const b200np_fab* const rho_box[2], double ro_0, double* const gp[2],
                                            const b200np_fab* const gp_box[2], double* const p_nd[2],
                                            const b200np_fab* const p_box[2], const double* inflow_vel0, double scaling_factor,
                                            int incremental, int proj_for_small_dt, double rtol, double atol, b200np_stats* stats)
{
    b200np_stats local{};
    b200np_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof(*st));
    if (!C || !velocity || !vel_box || !gp || !gp_box || !p_nd || !p_box) return st->status = B200NP_ERR_BAD_ARG;
    for (int l = 0; l < 2; ++l)
        if (!velocity[l] || !vel_box[l] || !gp[l] || !gp_box[l] || !p_nd[l] || !p_box[l]) return st->status = B200NP_ERR_BAD_ARG;
    const int use_old = (incremental || proj_for_small_dt);
    if (use_old && (!velocity_o || !velocity_o[0] || !velocity_o[1])) return st->status = B200NP_ERR_BAD_ARG;
    const bool var = density && (density[0] || density[1]);
    if (var && (!density[0] || !density[1] || !rho_box || !rho_box[0] || !rho_box[1])) return st->status = B200NP_ERR_BAD_ARG;
    b200np* hh[2] = {C->h0, C->h1};
    try {
        CK(cudaSetDevice(hh[0]->device));
        const Lev& g = hh[0]->lv[0].g;
        const int clo[3] = {0, 0, 0}, chi[3] = {g.n[0] - 1, g.n[1] - 1, g.n[2] - 1};
        if (!vel_box_ok(hh[0], vel_box[0]) || !box_covers(gp_box[0], clo, chi, 3)) return st->status = B200NP_ERR_BAD_ARG;
        if (!fine_box_ok(C, vel_box[1], 3, 1, false) || !fine_box_ok(C, gp_box[1], 3, 0, false)) return st->status = B200NP_ERR_BAD_ARG;
        if (var && (!box_covers(rho_box[0], clo, chi, 1) || !fine_box_ok(C, rho_box[1], 1, 0, false))) return st->status = B200NP_ERR_BAD_ARG;
        if (!nodal_box_ok(hh[0], p_box[0]) || !fine_box_ok(C, p_box[1], 1, 0, true)) return st->status = B200NP_ERR_BAD_ARG;
        hh[0]->launches = hh[1]->launches = 0;
        cudaStream_t stream = hh[0]->stream;
        CK(cudaEventRecord(hh[0]->ev[0], stream));
        const int set_inflow = (!proj_for_small_dt && !incremental);   // :81
        bool sv[2], so[2], sr[2], sg[2], sp[2], sin0;
        Fab fvel[2], fvelo[2], frho[2], fgp[2], fp[2];
        for (int l = 0; l < 2; ++l) {
            b200np* h = hh[l];
            double* dvel = stage_in(h, 0, velocity[l], vel_box[l], true, &sv[l], st);
            double* dvelo = stage_in(h, 1, use_old ? velocity_o[l] : nullptr, vel_box[l], true, &so[l], st);
            double* drho = stage_in(h, 2, var ? density[l] : nullptr, var ? rho_box[l] : nullptr, true, &sr[l], st);
            double* dgp = stage_in(h, 4, gp[l], gp_box[l], true, &sg[l], st);
            double* dp = stage_in(h, 5, p_nd[l], p_box[l], incremental != 0, &sp[l], st);
            if (l == 0) {
                fvel[l] = make_fab(dvel, vel_box[l]); fvelo[l] = make_fab(dvelo, vel_box[l]); frho[l] = make_fab(drho, var ? rho_box[l] : nullptr);
                fgp[l] = make_fab(dgp, gp_box[l]); fp[l] = make_fab(dp, p_box[l]);
            } else {
                fvel[l] = make_fab_fine(C, dvel, vel_box[l]); fvelo[l] = make_fab_fine(C, dvelo, vel_box[l]);
                frho[l] = make_fab_fine(C, drho, var ? rho_box[l] : nullptr);
                fgp[l] = make_fab_fine(C, dgp, gp_box[l]); fp[l] = make_fab_fine(C, dp, p_box[l]);
            }
        }
        double* din = stage_in(hh[0], 6, set_inflow ? inflow_vel0 : nullptr, vel_box[0], true, &sin0, st);
        Fab fin = make_fab(din, vel_box[0]);
        CK(cudaEventRecord(hh[0]->ev[4], stream));
        // per level: u += s gp / rho, sigma = s / rho (:39-71, :101-121); vel.setBndry(0) + inflow fill on the coarse level (:137-163)
        for (int l = 0; l < 2; ++l) {
            LevelData& L = hh[l]->lv[0];
            if (!incremental || use_old || var)
                LAUNCH(hh[0], k_pre_add_sigma, L.gc, 256, L.g, fvel[l], fgp[l], frho[l], fvelo[l], scaling_factor, ro_0, incremental ? 0 : 1,
                       use_old, var ? L.sigma : (double*)nullptr);
            long long total = (long long)fvel[l].nx * fvel[l].ny * fvel[l].nz;
            int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
            LAUNCH(hh[0], k_set_vel_ghosts, blocks, 256, L.g, fvel[l], l == 0 ? fin : Fab{}, l == 0 ? set_inflow : 0);
        }
        comp_set_sigma(C, var, scaling_factor / ro_0);
        // :181-266 (gp, p_nd copy-out and average_down(gp) included)
        int status = comp_core(C, fvel[0], fvel[1], fvelo[0], fvelo[1], use_old, fgp[0], fgp[1], incremental, fp[0], fp[1], incremental,
                               rtol, atol, st);
        CK(cudaEventRecord(hh[0]->ev[5], stream));
        for (int l = 0; l < 2; ++l) {
            stage_out(hh[l], 0, velocity[l], vel_box[l], sv[l], st);
            stage_out(hh[l], 4, gp[l], gp_box[l], sg[l], st);
            stage_out(hh[l], 5, p_nd[l], p_box[l], sp[l], st);
        }
        finish_stats(hh[0], st);
        st->launches = hh[0]->launches + hh[1]->launches;
        float ms;
        CK(cudaEventElapsedTime(&ms, hh[0]->ev[0], hh[0]->ev[4])); st->ms_h2d = ms;
        CK(cudaEventElapsedTime(&ms, hh[0]->ev[5], hh[0]->ev[1])); st->ms_d2h = ms;
        return st->status = status;
    } catch (int e) { return st->status = e; }
}

// incflo::ApplyNodalProjection with finest_level = 1 over multi-box MultiFabs: both levels' fabs are gathered into one array per field
// and level (level 1: the fine box in fine index space), the single-box composite path runs, the results are scattered back.
int b200np_composite_apply_nodal_projection_mf(b200np_composite_t* C, const b200np_mfab* const velocity[2], const b200np_mfab* const velocity_o[2],
                                               const b200np_mfab* const density[2], double ro_0, const b200np_mfab* const gp[2],
                                               const b200np_mfab* const p_nd[2], const b200np_mfab* inflow_vel0, double scaling_factor,
                                               int incremental, int proj_for_small_dt, double rtol, double atol, b200np_stats* stats)
{
    b200np_stats local{};
    b200np_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof(*st));
    if (!C || !velocity || !gp || !p_nd) return st->status = B200NP_ERR_BAD_ARG;
    const int use_old = (incremental || proj_for_small_dt);
    const bool var = density && (density[0] || density[1]);
    if (var && (!density[0] || !density[1])) return st->status = B200NP_ERR_BAD_ARG;
    if (use_old && (!velocity_o || !velocity_o[0] || !velocity_o[1])) return st->status = B200NP_ERR_BAD_ARG;
    b200np* hh[2] = {C->h0, C->h1};
    try {
        CK(cudaSetDevice(hh[0]->device));
        // per level: cells, cells grown by one, nodes -- level 0 over the domain, level 1 over the fine box (fine index space)
        b200np_fab cells[2]{}, grown[2]{}, nodes[2]{};
        mf_slab_boxes(hh[0], cells[0], grown[0], nodes[0]);
        for (int d = 0; d < 3; ++d) {
            cells[1].lo[d] = 2 * C->box.lo[d]; cells[1].hi[d] = 2 * C->box.hi[d] + 1;
            grown[1].lo[d] = cells[1].lo[d] - 1; grown[1].hi[d] = cells[1].hi[d] + 1;
            nodes[1].lo[d] = cells[1].lo[d]; nodes[1].hi[d] = cells[1].hi[d] + 1;
        }
        cells[1].ncomp = 1; grown[1].ncomp = 3; nodes[1].ncomp = 1;
        const int set_inflow = (!proj_for_small_dt && !incremental);
        for (int l = 0; l < 2; ++l) {
            if (!mf_boxes_ok(velocity[l], cells[l].lo, cells[l].hi, 3, true) || velocity[l]->ngrow < 1) return st->status = B200NP_ERR_BAD_ARG;
            if (use_old && !mf_boxes_ok(velocity_o[l], cells[l].lo, cells[l].hi, 3, true)) return st->status = B200NP_ERR_BAD_ARG;
            if (var && !mf_boxes_ok(density[l], cells[l].lo, cells[l].hi, 1, true)) return st->status = B200NP_ERR_BAD_ARG;
            if (!mf_boxes_ok(gp[l], cells[l].lo, cells[l].hi, 3, true) || !mf_boxes_ok(p_nd[l], nodes[l].lo, nodes[l].hi, 1, false)) return st->status = B200NP_ERR_BAD_ARG;
        }
        if (inflow_vel0 && set_inflow && (!mf_boxes_ok(inflow_vel0, cells[0].lo, cells[0].hi, 3, true) || inflow_vel0->ngrow < 1)) return st->status = B200NP_ERR_BAD_ARG;
        mf_events(hh[0]);
        b200np_stats outer{};
        CK(cudaEventRecord(hh[0]->mf_ev[0], hh[0]->stream));
        MfField Fv[2], Fo[2], Fr[2], Fg[2], Fp[2], Fi;
        double *svel[2], *svelo[2], *srho[2], *sgp[2], *sp[2], *sin = nullptr;
        b200np_fab gbox[2];
        long long extra = 0;
        for (int l = 0; l < 2; ++l) {
            b200np* h = hh[l];
            mf_map(h, 0, velocity[l], 3, true, &outer, Fv[l]);
            mf_map(h, 1, use_old ? velocity_o[l] : nullptr, 3, true, &outer, Fo[l]);
            mf_map(h, 2, var ? density[l] : nullptr, 1, true, &outer, Fr[l]);
            mf_map(h, 3, gp[l], 3, true, &outer, Fg[l]);
            mf_map(h, 4, p_nd[l], 1, incremental != 0, &outer, Fp[l]);
            gbox[l] = cells[l]; gbox[l].ncomp = 3;
            svel[l] = mf_slab(h, 0, grown[l]);
            svelo[l] = Fo[l].mf ? mf_slab(h, 1, grown[l]) : nullptr;
            srho[l] = Fr[l].mf ? mf_slab(h, 2, cells[l]) : nullptr;
            sgp[l] = mf_slab(h, 3, gbox[l]);
            sp[l] = mf_slab(h, 4, nodes[l]);
            h->launches = 0;
            CK(cudaMemsetAsync(svel[l], 0, fab_bytes(&grown[l]), h->stream));   // ghost cells: set by setBndry(0) / the inflow fill
            mf_gather(h, Fv[l], 3, svel[l], grown[l], MF_VALID);
            if (Fo[l].mf) { CK(cudaMemsetAsync(svelo[l], 0, fab_bytes(&grown[l]), h->stream)); mf_gather(h, Fo[l], 3, svelo[l], grown[l], MF_VALID); }
            mf_gather(h, Fr[l], 1, srho[l], cells[l], MF_VALID);
            mf_gather(h, Fg[l], 3, sgp[l], gbox[l], MF_VALID);
            if (incremental) mf_gather(h, Fp[l], 1, sp[l], nodes[l], MF_VALID);
            if (l == 0) {
                mf_map(h, 5, (inflow_vel0 && set_inflow) ? inflow_vel0 : nullptr, 3, true, &outer, Fi);
                if (Fi.mf) {
                    sin = mf_slab(h, 5, grown[0]);
                    CK(cudaMemsetAsync(sin, 0, fab_bytes(&grown[0]), h->stream));
                    mf_gather(h, Fi, 3, sin, grown[0], MF_VALID_BC);
                }
            }
            extra += h->launches;
        }
        double* const a_vel[2] = {svel[0], svel[1]};
        const double* const a_velo[2] = {svelo[0], svelo[1]};
        const double* const a_rho[2] = {srho[0], srho[1]};
        double* const a_gp[2] = {sgp[0], sgp[1]};
        double* const a_p[2] = {sp[0], sp[1]};
        const b200np_fab* const b_vel[2] = {&grown[0], &grown[1]};
        const b200np_fab* const b_rho[2] = {&cells[0], &cells[1]};
        const b200np_fab* const b_gp[2] = {&gbox[0], &gbox[1]};
        const b200np_fab* const b_p[2] = {&nodes[0], &nodes[1]};
        int rc = b200np_composite_apply_nodal_projection(C, a_vel, b_vel, use_old ? a_velo : nullptr, var ? a_rho : nullptr, var ? b_rho : nullptr, ro_0,
                                                         a_gp, b_gp, a_p, b_p, sin, scaling_factor, incremental, proj_for_small_dt, rtol, atol, st);
        if (rc != B200NP_OK && rc != B200NP_ERR_NOT_CONVERGED && rc != B200NP_ERR_DIVERGED) return rc;
        for (int l = 0; l < 2; ++l) {
            b200np* h = hh[l];
            h->launches = 0;
            mf_scatter(h, Fv[l], 3, svel[l], grown[l], MF_VALID_BC);
            mf_scatter(h, Fg[l], 3, sgp[l], gbox[l], MF_VALID);
            mf_scatter(h, Fp[l], 1, sp[l], nodes[l], MF_VALID);
            extra += h->launches;
            mf_copy_back(h, 0, Fv[l], &outer); mf_copy_back(h, 3, Fg[l], &outer); mf_copy_back(h, 4, Fp[l], &outer);
        }
        CK(cudaEventRecord(hh[0]->mf_ev[1], hh[0]->stream));
        CK(cudaEventSynchronize(hh[0]->mf_ev[1]));
        float ms;
        CK(cudaEventElapsedTime(&ms, hh[0]->mf_ev[0], hh[0]->mf_ev[1])); st->ms_total = ms;
        st->launches += extra; st->h2d_bytes += outer.h2d_bytes; st->d2h_bytes += outer.d2h_bytes;
        return st->status = rc;
    } catch (int e) { return st->status = e; }
}

// ---- test hooks ---------------------------------------------------------------------------
static double* level_array(b200np* h, int lev, int which)
{
    LevelData& L = h->lv[lev];
    switch (which) {
    case B200NP_A_SOL: return L.sol;
    case B200NP_A_RHS: return L.rhs;
    case B200NP_A_RES: return L.res;
    case B200NP_A_COR: return L.cor;
    case B200NP_A_RESCOR: return L.rescor;
    case B200NP_A_SIGMA: return L.sigma;
    default: return nullptr;
    }
}

int b200np_level_set(b200np_t* h, int lev, int which, const double* host)
{
    if (!h || lev < 0 || lev >= (int)h->lv.size() || !host) return B200NP_ERR_BAD_ARG;
    double* d = level_array(h, lev, which);
    if (!d) return B200NP_ERR_BAD_ARG;
    const Lev& g = h->lv[lev].g;
    try {
        CK(cudaSetDevice(h->device));
        if (which == B200NP_A_SIGMA)
            CK(cudaMemcpy2DAsync(d, g.cpx * sizeof(double), host, g.n[0] * sizeof(double), g.n[0] * sizeof(double),
                                 (size_t)g.n[1] * g.cnzl, cudaMemcpyHostToDevice, h->stream));
        else
            CK(cudaMemcpy2DAsync(d, g.px * sizeof(double), host, g.nn[0] * sizeof(double), g.nn[0] * sizeof(double),
                                 (size_t)g.nn[1] * g.nzl, cudaMemcpyHostToDevice, h->stream));
        // the solver stream is non-blocking: order the copy on it (a pageable cudaMemcpy on the
        // legacy stream may still be in flight when the next kernel starts)
        CK(cudaStreamSynchronize(h->stream));
    } catch (int e) { return e; }
    return B200NP_OK;
}

int b200np_level_get(b200np_t* h, int lev, int which, double* host)
{
    if (!h || lev < 0 || lev >= (int)h->lv.size() || !host) return B200NP_ERR_BAD_ARG;
    double* d = level_array(h, lev, which);
    if (!d) return B200NP_ERR_BAD_ARG;
    const Lev& g = h->lv[lev].g;
    try {
        CK(cudaSetDevice(h->device));
        CK(cudaStreamSynchronize(h->stream));
        if (which == B200NP_A_SIGMA)
            CK(cudaMemcpy2D(host, g.n[0] * sizeof(double), d, g.cpx * sizeof(double), g.n[0] * sizeof(double),
                            (size_t)g.n[1] * g.cnzl, cudaMemcpyDeviceToHost));
        else
            CK(cudaMemcpy2D(host, g.nn[0] * sizeof(double), d, g.px * sizeof(double), g.nn[0] * sizeof(double),
                            (size_t)g.nn[1] * g.nzl, cudaMemcpyDeviceToHost));
    } catch (int e) { return e; }
    return B200NP_OK;
}

static int run_op(b200np* h, int lev, int op, int arg)
{
    const int nl = (int)h->lv.size();
    LevelData& L = h->lv[lev];
    switch (op) {
    case B200NP_OP_SMOOTH: {
        double *x = L.cor, *y = L.cor2;
        // arg bit 16: "cor is zero" smooth call exactly as the V-cycle's pre-smooth issues it: either the first sweep
        // never reads cor (zero-start kernels), or cor is cleared first
        const bool zs = (arg & B200NP_SMOOTH_ZERO_START) != 0;
        if (zs && !(h->zero_start && h->smoother_version >= 3 && L.iso))
            CK(cudaMemsetAsync(L.cor - L.g.ps, 0, (size_t)L.g.ps * (L.g.nzl + 2) * sizeof(double), h->stream));
        smooth_sweeps(h, L, x, y, L.res, arg & 0xffff, zs);
        if (x != L.cor) CK(cudaMemcpyAsync(L.cor, x, (size_t)L.g.ps * L.g.nzl * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        break;
    }
    case B200NP_OP_RESIDUAL: residual(h, L, L.cor, L.res, L.rescor, nullptr); break;
    case B200NP_OP_RESTRICT: if (lev + 1 >= nl) return B200NP_ERR_BAD_ARG; restrict_to(h, lev); break;
    case B200NP_OP_INTERP: if (lev + 1 >= nl) return B200NP_ERR_BAD_ARG; interp_add(h, lev); break;
    case B200NP_OP_BOTTOM: bottom_solve(h); break;
    case B200NP_OP_VCYCLE: vcycle_launch(h, 0); break;
    case B200NP_OP_COARSEN_SIGMA: coarsen_sigma(h); break;
    default: return B200NP_ERR_BAD_ARG;
    }
    return B200NP_OK;
}

int b200np_level_op(b200np_t* h, int lev, int op, int arg)
{
    if (!h || lev < 0 || lev >= (int)h->lv.size()) return B200NP_ERR_BAD_ARG;
    try {
        CK(cudaSetDevice(h->device));
        int rc = run_op(h, lev, op, arg);
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaGetLastError());
        return rc;
    } catch (int e) { return e; }
}

int b200np_time_op(b200np_t* h, int lev, int op, int arg, int reps, double* ms)
{
    if (!h || lev < 0 || lev >= (int)h->lv.size() || reps < 1 || !ms) return B200NP_ERR_BAD_ARG;
    try {
        CK(cudaSetDevice(h->device));
        int rc = run_op(h, lev, op, arg);  // warm-up
        if (rc) return rc;
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaEventRecord(h->ev[0], h->stream));
        for (int r = 0; r < reps; ++r) run_op(h, lev, op, arg);
        CK(cudaEventRecord(h->ev[1], h->stream));
        CK(cudaEventSynchronize(h->ev[1]));
        float t;
        CK(cudaEventElapsedTime(&t, h->ev[0], h->ev[1]));
        *ms = t / reps;
        return B200NP_OK;
    } catch (int e) { return e; }
}

}  // extern "C"
