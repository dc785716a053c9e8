// shim_check.cpp -- exercises include/B200NodalProjector.H (the C++ mirror of Hydro::NodalProjector /
// incflo::ApplyNodalProjection) the way incflo_apply_nodal_projection.cpp:181-219 uses the reference.
//   shim_check host                         : host-logic checks, no GPU needed (BC map, options, abort behaviour)
//   shim_check project in.bin out.bin nx ny nz dx bclo(3) bchi(3) var
//       in.bin : vel (3,(nz+2),(ny+2),(nx+2)) [+ sigma (nz,ny,nx) if var]; out.bin: vel, phi, gphi, iters
//   shim_check composite in.bin out.bin nx ny nz dx bclo(3) bchi(3) flo(3) fhi(3)
//       two AMR levels, constant sigma 0.37, fine box = coarse cells [flo, fhi] refined by 2, 1 ghost cell per level;
//       in.bin: vel0, vel1; out.bin: vel0, vel1, phi0, phi1, gphi0, gphi1, iters
//   shim_check mac n                      : the MacProjector call sequence of incflo_compute_MAC_projected_velocities.cpp:96-132,
//       :287-298 (constant beta, then face arrays, periodic x/y + walls z); checks that the projected face velocity is
//       discretely divergence-free and that project(mac_phi, ..) started from the converged phi does nothing
//   shim_check multibox n max_grid        : incflo::ApplyNodalProjection over a MultiFab of max_grid^3 boxes (ng = 2) against
//       the same call on one box -- must agree bit for bit (periodic x/y, walls z, variable density)
//   shim_check eb in.bin out.bin nx ny nz dx bclo(3) bchi(3) sigma ebflow
//       the AMREX_USE_EB call sequence of incflo_apply_nodal_projection.cpp:130-136, :181-219 through b200::EBNodalProjector:
//       in.bin: vel (3, nz+2, ny+2, nx+2), vfrac, intg (18), bnorm (3), bintg (8); constant sigma; ebflow != 0: set_eb_velocity with
//       eb_flow.vel_mag = ebflow, then getLinOp().setEBInflowVelocity; out.bin: vel, phi, gphi, eb_vel (3, with 1 ghost cell), iters
#include "../../include/B200NodalProjector.H"
#include "../../include/B200MacProjector.H"
#include "../../include/B200EBNodalProjector.H"

#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>

using namespace b200;

static void throwing_abort(const char* msg) { throw std::runtime_error(msg); }

#define EXPECT(cond)                                                              \
    do {                                                                          \
        if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } \
    } while (0)

static int host_checks()
{
    abort_handler() = throwing_abort;
    // incflo_projection_bc.cpp:5-41
    std::array<bool, 3> per{true, false, false};
    auto r = get_projection_bc(per, {BC::undefined, BC::pressure_outflow, BC::no_slip_wall});
    EXPECT(r[0] == LinOpBCType::Periodic && r[1] == LinOpBCType::Dirichlet && r[2] == LinOpBCType::Neumann);
    r = get_projection_bc({false, false, false}, {BC::mass_inflow, BC::direction_dependent, BC::mixed});
    EXPECT(r[0] == LinOpBCType::inflow && r[1] == LinOpBCType::inflow && r[2] == LinOpBCType::inflow);
    r = get_projection_bc({false, false, false}, {BC::pressure_inflow, BC::slip_wall, BC::slip_wall});
    EXPECT(r[0] == LinOpBCType::Dirichlet && r[1] == LinOpBCType::Neumann);
    bool threw = false;
    try { get_projection_bc({false, true, true}, {BC::undefined, BC::undefined, BC::undefined}); }
    catch (const std::runtime_error& e) { threw = std::string(e.what()) == "get_projection_bc: undefined BC type"; }
    EXPECT(threw);
    // nodal_proj.* keys and defaults (src/incflo.H:436-458)
    b200np_opts o = nodal_proj_options();
    EXPECT(o.maxiter == 100 && o.bottom_maxiter == 100 && o.bottom_rtol == 1e-4 && o.mg_max_coarsening_level == 100);
    EXPECT(o.num_pre_smooth == 2 && o.num_post_smooth == 2 && o.verbose == 0);
    o = nodal_proj_options({{"verbose", "2"}, {"maxiter", "7"}, {"bottom_solver", "smoother"}, {"mg_rtol", "1e-9"}});
    EXPECT(o.verbose == 2 && o.maxiter == 7 && o.bottom_solver == 1);
    threw = false;
    try { nodal_proj_options({{"no_such_key", "1"}}); } catch (const std::runtime_error&) { threw = true; }
    EXPECT(threw);
    threw = false;
    try { nodal_proj_options({{"bottom_solver", "hypre"}}); } catch (const std::runtime_error&) { threw = true; }
    EXPECT(threw);
    // project() before setDomainBC aborts; multi-level aborts
    const int n[3] = {8, 8, 8};
    std::vector<double> v((size_t)3 * 10 * 10 * 10, 0.0);
    Geometry g{{8, 8, 8}, {0.125, 0.125, 0.125}, {true, true, true}};
    {
        NodalProjector np({Fab::make(v.data(), n, 1, 3)}, 1.0, {g}, LPInfo().setMaxCoarseningLevel(100));
        threw = false;
        try { np.project(1e-11, 1e-14); } catch (const std::runtime_error&) { threw = true; }
        EXPECT(threw);
    }
    // two levels: level 1 must be one box at ratio 2 aligned with the coarse cells; three levels abort
    threw = false;
    try { NodalProjector np2({Fab::make(v.data(), n, 1, 3), Fab::make(v.data(), n, 1, 3)}, 1.0, {g, g}); }   // geom[1] is not the refined geometry
    catch (const std::runtime_error&) { threw = true; }
    EXPECT(threw);
    threw = false;
    try { NodalProjector np3({Fab::make(v.data(), n, 1, 3), Fab::make(v.data(), n, 1, 3), Fab::make(v.data(), n, 1, 3)}, 1.0, {g, g, g}); }
    catch (const std::runtime_error&) { threw = true; }
    EXPECT(threw);
    {
        Geometry g1{{16, 16, 16}, {0.0625, 0.0625, 0.0625}, {true, true, true}};
        const int vlo[3] = {4, 4, 4}, vhi[3] = {11, 11, 11};   // fine cells 4..11 = coarse cells 2..5
        NodalProjector np4({Fab::make(v.data(), n, 1, 3), Fab::make_box(v.data(), vlo, vhi, 1, 3)}, 1.0, {g, g1});
        const int odd[3] = {5, 4, 4};
        threw = false;
        try { NodalProjector np5({Fab::make(v.data(), n, 1, 3), Fab::make_box(v.data(), odd, vhi, 1, 3)}, 1.0, {g, g1}); }
        catch (const std::runtime_error&) { threw = true; }
        EXPECT(threw);
    }
    // Fab geometry: ld.velocity (ng=3), ld.p_nd (nodal, ng=0)   src/setup/incflo_arrays.cpp:9-26
    Fab f = Fab::make(nullptr, n, 3, 3);
    EXPECT(f.box.lo[0] == -3 && f.box.hi[2] == 10 && f.size() == (size_t)3 * 14 * 14 * 14);
    Fab pn = Fab::make(nullptr, n, 0, 1, true);
    EXPECT(pn.box.lo[1] == 0 && pn.box.hi[1] == 8 && pn.size() == (size_t)9 * 9 * 9);
    // MultiFab view: what mfab_of(amrex::MultiFab&) builds
    {
        std::vector<double> a(1000), b(1000);
        const int lo0[3] = {0, 0, 0}, hi0[3] = {3, 7, 7}, lo1[3] = {4, 0, 0}, hi1[3] = {7, 7, 7};
        MultiFab mf({Fab::make_box(a.data(), lo0, hi0, 1, 1), Fab::make_box(b.data(), lo1, hi1, 1, 1)}, 1, 1);
        const b200np_mfab* c = mf.c();
        EXPECT(c && c->nfabs == 2 && c->ngrow == 1 && c->ncomp == 1 && c->data[1] == b.data());
        EXPECT(c->box[1].lo[0] == 3 && c->box[1].hi[0] == 8 && c->box[0].lo[2] == -1);
        EXPECT(MultiFab().c() == nullptr);
    }
    std::printf("shim host checks OK\n");
    return 0;
}

// incflo::ApplyNodalProjection on a multi-box LevelData against the single-box call
static int multibox(int argc, char** argv)
{
    if (argc < 4) { std::printf("usage: shim_check multibox n max_grid\n"); return 2; }
    abort_handler() = throwing_abort;
    const int N = std::atoi(argv[2]), mg = std::atoi(argv[3]), ng = 2;
    const int n[3] = {N, N, N};
    Geometry g{{N, N, N}, {1.0 / N, 1.0 / N, 1.0 / N}, {true, true, false}};
    std::array<LinOpBCType, 3> lo{LinOpBCType::Periodic, LinOpBCType::Periodic, LinOpBCType::Neumann}, hi = lo;
    const int S = N + 2 * ng;
    std::vector<double> vel((size_t)3 * S * S * S), rho((size_t)S * S * S), gp((size_t)3 * N * N * N, 0.0), p((size_t)(N + 1) * (N + 1) * (N + 1), 0.0);
    unsigned long long seed = 88172645463325252ull;
    auto rnd = [&]() { seed ^= seed << 13; seed ^= seed >> 7; seed ^= seed << 17; return (double)(seed % 2000001ull) / 1.0e6 - 1.0; };
    for (auto& x : vel) x = rnd();
    for (auto& x : rho) x = 1.0 + 0.5 * (rnd() + 1.0);
    for (auto& x : gp) x = 0.1 * rnd();
    try {
        IncfloNodalProjection proj(g, lo, hi);
        // (a) one box
        std::vector<double> v1 = vel, g1 = gp, p1 = p;
        Fab fv = Fab::make(v1.data(), n, ng, 3), fr = Fab::make(rho.data(), n, ng, 1), fg = Fab::make(g1.data(), n, 0, 3), fp = Fab::make(p1.data(), n, 0, 1, true);
        proj.ApplyNodalProjection(&fr, 1.0, fv, nullptr, fg, fp, nullptr, 0.01, false);
        const int it1 = proj.stats().iters;
        // (b) boxes of mg^3 cells, each fab with its own ghost frame (what MFIter hands out)
        struct Store { std::vector<std::vector<double>> d; };
        Store sv, sr, sg, sp;
        auto chop = [&](const std::vector<double>& full, int ncomp, int ngr, bool nodal, Store& st) {
            std::vector<Fab> fabs;
            const int e = nodal ? 1 : 0, FS = N + 2 * ngr + e;
            for (int k0 = 0; k0 < N; k0 += mg) for (int j0 = 0; j0 < N; j0 += mg) for (int i0 = 0; i0 < N; i0 += mg) {
                const int vlo[3] = {i0, j0, k0}, vhi[3] = {std::min(i0 + mg, N) - 1, std::min(j0 + mg, N) - 1, std::min(k0 + mg, N) - 1};
                const int bx = vhi[0] - vlo[0] + 1 + 2 * ngr + e, by = vhi[1] - vlo[1] + 1 + 2 * ngr + e, bz = vhi[2] - vlo[2] + 1 + 2 * ngr + e;
                st.d.emplace_back((size_t)ncomp * bx * by * bz);
                auto& b = st.d.back();
                for (int c = 0; c < ncomp; ++c) for (int k = 0; k < bz; ++k) for (int j = 0; j < by; ++j) for (int i = 0; i < bx; ++i)
                    b[(((size_t)c * bz + k) * by + j) * bx + i] = full[(((size_t)c * FS + (k0 + k)) * FS + (j0 + j)) * FS + (i0 + i)];
                fabs.push_back(Fab::make_box(b.data(), vlo, vhi, ngr, ncomp, nodal));
            }
            return MultiFab(std::move(fabs), ngr, ncomp);
        };
        MultiFab mv = chop(vel, 3, ng, false, sv), mr = chop(rho, 1, ng, false, sr), mgp = chop(gp, 3, 0, false, sg), mp = chop(p, 1, 0, true, sp);
        proj.ApplyNodalProjection(&mr, 1.0, mv, nullptr, mgp, mp, nullptr, 0.01, false);
        if (proj.stats().iters != it1) { std::printf("iterations differ: %d vs %d\n", proj.stats().iters, it1); return 5; }
        // compare the valid regions
        size_t bad = 0, f = 0;
        for (int k0 = 0; k0 < N; k0 += mg) for (int j0 = 0; j0 < N; j0 += mg) for (int i0 = 0; i0 < N; i0 += mg, ++f) {
            const Fab& a = mv.fabs[f];
            const int bx = a.box.hi[0] - a.box.lo[0] + 1, by = a.box.hi[1] - a.box.lo[1] + 1, bz = a.box.hi[2] - a.box.lo[2] + 1;
            for (int c = 0; c < 3; ++c) for (int k = ng; k < bz - ng; ++k) for (int j = ng; j < by - ng; ++j) for (int i = ng; i < bx - ng; ++i)
                if (a.p[(((size_t)c * bz + k) * by + j) * bx + i] != v1[(((size_t)c * S + (k0 + k)) * S + (j0 + j)) * S + (i0 + i)]) ++bad;
            const Fab& q = mp.fabs[f];
            const int qx = q.box.hi[0] - q.box.lo[0] + 1, qy = q.box.hi[1] - q.box.lo[1] + 1, qz = q.box.hi[2] - q.box.lo[2] + 1;
            for (int k = 0; k < qz; ++k) for (int j = 0; j < qy; ++j) for (int i = 0; i < qx; ++i)
                if (q.p[((size_t)k * qy + j) * qx + i] != p1[((size_t)(k0 + k) * (N + 1) + (j0 + j)) * (N + 1) + (i0 + i)]) ++bad;
        }
        if (bad) { std::printf("%zu values differ between the multi-box and the single-box call\n", bad); return 6; }
        std::printf("shim multibox OK: %zu boxes, %d V-cycles\n", mv.fabs.size(), it1);
    } catch (const std::runtime_error& e) {
        std::printf("amrex::Abort::%s\n", e.what());
        return 1;
    }
    return 0;
}

// incflo::ApplyNodalProjection with finest_level = 1 through IncfloCompositeProjection: one box per level against both levels chopped
// into boxes of mg cells (level 1 in fine index space) -- bit-identical
static int composite_mf(int argc, char** argv)
{
    if (argc < 4) { std::printf("usage: shim_check composite_mf n max_grid\n"); return 2; }
    abort_handler() = throwing_abort;
    const int N = std::atoi(argv[2]), mg = std::atoi(argv[3]), ng = 2;
    Geometry g{{N, N, N}, {1.0 / N, 1.0 / N, 1.0 / N}, {true, true, false}};
    std::array<LinOpBCType, 3> lo{LinOpBCType::Periodic, LinOpBCType::Periodic, LinOpBCType::Neumann}, hi = lo;
    const std::array<int, 3> flo{N / 4, N / 4, N / 4}, fhi{3 * N / 4 - 1, 3 * N / 4 - 1, 3 * N / 4 - 1};
    const int nl[2][3] = {{N, N, N}, {N, N, N}};                 // cells per level (level 1: 2 * N / 2)
    const int org[2][3] = {{0, 0, 0}, {2 * flo[0], 2 * flo[1], 2 * flo[2]}};
    unsigned long long seed = 1234567891234567ull;
    auto rnd = [&]() { seed ^= seed << 13; seed ^= seed >> 7; seed ^= seed << 17; return (double)(seed % 2000001ull) / 1.0e6 - 1.0; };
    std::vector<double> vel[2], rho[2], gp[2], p[2];
    for (int l = 0; l < 2; ++l) {
        const int S = N + 2 * ng;
        vel[l].resize((size_t)3 * S * S * S); rho[l].resize((size_t)S * S * S); gp[l].resize((size_t)3 * N * N * N); p[l].assign((size_t)(N + 1) * (N + 1) * (N + 1), 0.0);
        for (auto& x : vel[l]) x = rnd();
        for (auto& x : rho[l]) x = 1.0 + 0.5 * (rnd() + 1.0);
        for (auto& x : gp[l]) x = 0.1 * rnd();
    }
    try {
        IncfloCompositeProjection proj(g, lo, hi, flo, fhi);
        // (a) one box per level
        std::vector<double> v1[2] = {vel[0], vel[1]}, g1[2] = {gp[0], gp[1]}, p1[2] = {p[0], p[1]};
        auto box = [&](double* d, int l, int ngr, int nc, bool nodal) {
            const int vlo[3] = {org[l][0], org[l][1], org[l][2]}, vhi[3] = {org[l][0] + nl[l][0] - 1, org[l][1] + nl[l][1] - 1, org[l][2] + nl[l][2] - 1};
            return Fab::make_box(d, vlo, vhi, ngr, nc, nodal);
        };
        std::array<Fab, 2> fv{box(v1[0].data(), 0, ng, 3, false), box(v1[1].data(), 1, ng, 3, false)};
        std::array<Fab, 2> fr{box(rho[0].data(), 0, ng, 1, false), box(rho[1].data(), 1, ng, 1, false)};
        std::array<Fab, 2> fg{box(g1[0].data(), 0, 0, 3, false), box(g1[1].data(), 1, 0, 3, false)};
        std::array<Fab, 2> fp{box(p1[0].data(), 0, 0, 1, true), box(p1[1].data(), 1, 0, 1, true)};
        proj.ApplyNodalProjection(&fr, 1.0, fv, nullptr, fg, fp, nullptr, 0.01, false);
        const int it1 = proj.stats().iters;
        // (b) boxes of mg^3 cells on both levels
        std::vector<std::vector<double>> store;
        auto chop = [&](const std::vector<double>& full, int l, int ncomp, int ngr, bool nodal) {
            std::vector<Fab> fabs;
            const int e = nodal ? 1 : 0, FS = N + 2 * ngr + e;
            for (int k0 = 0; k0 < N; k0 += mg) for (int j0 = 0; j0 < N; j0 += mg) for (int i0 = 0; i0 < N; i0 += mg) {
                const int vlo[3] = {org[l][0] + i0, org[l][1] + j0, org[l][2] + k0};
                const int vhi[3] = {org[l][0] + std::min(i0 + mg, N) - 1, org[l][1] + std::min(j0 + mg, N) - 1, org[l][2] + std::min(k0 + mg, N) - 1};
                const int bx = vhi[0] - vlo[0] + 1 + 2 * ngr + e, by = vhi[1] - vlo[1] + 1 + 2 * ngr + e, bz = vhi[2] - vlo[2] + 1 + 2 * ngr + e;
                store.emplace_back((size_t)ncomp * bx * by * bz);
                auto& b = store.back();
                for (int c = 0; c < ncomp; ++c) for (int k = 0; k < bz; ++k) for (int j = 0; j < by; ++j) for (int i = 0; i < bx; ++i)
                    b[(((size_t)c * bz + k) * by + j) * bx + i] = full[(((size_t)c * FS + (k0 + k)) * FS + (j0 + j)) * FS + (i0 + i)];
                fabs.push_back(Fab::make_box(b.data(), vlo, vhi, ngr, ncomp, nodal));
            }
            return MultiFab(std::move(fabs), ngr, ncomp);
        };
        store.reserve(8 * (size_t)((N + mg - 1) / mg) * ((N + mg - 1) / mg) * ((N + mg - 1) / mg) + 8);   // Fab pointers stay valid
        std::array<MultiFab, 2> mv{chop(vel[0], 0, 3, ng, false), chop(vel[1], 1, 3, ng, false)};
        std::array<MultiFab, 2> mr{chop(rho[0], 0, 1, ng, false), chop(rho[1], 1, 1, ng, false)};
        std::array<MultiFab, 2> mgp{chop(gp[0], 0, 3, 0, false), chop(gp[1], 1, 3, 0, false)};
        std::array<MultiFab, 2> mp{chop(p[0], 0, 1, 0, true), chop(p[1], 1, 1, 0, true)};
        proj.ApplyNodalProjection(&mr, 1.0, mv, nullptr, mgp, mp, nullptr, 0.01, false);
        if (proj.stats().iters != it1) { std::printf("iterations differ: %d vs %d\n", proj.stats().iters, it1); return 5; }
        size_t bad = 0;
        const int S = N + 2 * ng;
        for (int l = 0; l < 2; ++l) {
            size_t f = 0;
            for (int k0 = 0; k0 < N; k0 += mg) for (int j0 = 0; j0 < N; j0 += mg) for (int i0 = 0; i0 < N; i0 += mg, ++f) {
                const Fab& a = mv[l].fabs[f];
                const int bx = a.box.hi[0] - a.box.lo[0] + 1, by = a.box.hi[1] - a.box.lo[1] + 1, bz = a.box.hi[2] - a.box.lo[2] + 1;
                for (int c = 0; c < 3; ++c) for (int k = ng; k < bz - ng; ++k) for (int j = ng; j < by - ng; ++j) for (int i = ng; i < bx - ng; ++i)
                    if (a.p[(((size_t)c * bz + k) * by + j) * bx + i] != v1[l][(((size_t)c * S + (k0 + k)) * S + (j0 + j)) * S + (i0 + i)]) ++bad;
                const Fab& q = mp[l].fabs[f];
                const int qx = q.box.hi[0] - q.box.lo[0] + 1, qy = q.box.hi[1] - q.box.lo[1] + 1, qz = q.box.hi[2] - q.box.lo[2] + 1;
                for (int k = 0; k < qz; ++k) for (int j = 0; j < qy; ++j) for (int i = 0; i < qx; ++i)
                    if (q.p[((size_t)k * qy + j) * qx + i] != p1[l][((size_t)(k0 + k) * (N + 1) + (j0 + j)) * (N + 1) + (i0 + i)]) ++bad;
            }
        }
        if (bad) { std::printf("%zu values differ between the multi-box and the single-box call\n", bad); return 6; }
        std::printf("shim composite_mf OK: %zu + %zu boxes, %d iterations\n", mv[0].fabs.size(), mv[1].fabs.size(), it1);
    } catch (const std::runtime_error& e) {
        std::printf("amrex::Abort::%s\n", e.what());
        return 1;
    }
    return 0;
}

// the call sequence of incflo_apply_nodal_projection.cpp:181-219 with finest_level = 1
static int composite(int argc, char** argv)
{
    if (argc < 20) { std::printf("usage: see header comment\n"); return 2; }
    abort_handler() = throwing_abort;
    const int n[3] = {std::atoi(argv[4]), std::atoi(argv[5]), std::atoi(argv[6])};
    const double dx = std::atof(argv[7]);
    std::array<LinOpBCType, 3> lo, hi;
    Geometry g0, g1;
    int flo[3], fhi[3], vlo[3], vhi[3], nf[3];
    for (int d = 0; d < 3; ++d) {
        lo[d] = (LinOpBCType)std::atoi(argv[8 + d]); hi[d] = (LinOpBCType)std::atoi(argv[11 + d]);
        flo[d] = std::atoi(argv[14 + d]); fhi[d] = std::atoi(argv[17 + d]);
        g0.n_cell[d] = n[d]; g0.dx[d] = dx; g0.is_periodic[d] = lo[d] == LinOpBCType::Periodic;
        g1.n_cell[d] = 2 * n[d]; g1.dx[d] = 0.5 * dx; g1.is_periodic[d] = g0.is_periodic[d];
        vlo[d] = 2 * flo[d]; vhi[d] = 2 * fhi[d] + 1; nf[d] = vhi[d] - vlo[d] + 1;
    }
    const size_t nv0 = (size_t)3 * (n[0] + 2) * (n[1] + 2) * (n[2] + 2), nv1 = (size_t)3 * (nf[0] + 2) * (nf[1] + 2) * (nf[2] + 2);
    std::vector<double> vel0(nv0), vel1(nv1);
    std::ifstream in(argv[2], std::ios::binary);
    in.read((char*)vel0.data(), nv0 * 8);
    in.read((char*)vel1.data(), nv1 * 8);
    if (!in) { std::printf("short input\n"); return 3; }
    try {
        LPInfo info;
        info.setMaxCoarseningLevel(100);
        std::vector<Fab> velv{Fab::make(vel0.data(), n, 1, 3), Fab::make_box(vel1.data(), vlo, vhi, 1, 3)};
        auto nodal_projector = std::make_unique<NodalProjector>(velv, 0.37, std::vector<Geometry>{g0, g1}, info);
        nodal_projector->setDomainBC(lo, hi);
        nodal_projector->project(1e-11, 1e-14);
        auto phi = nodal_projector->getPhi();
        auto gradphi = nodal_projector->getGradPhi();
        if (phi.size() != 2 || gradphi.size() != 2) { std::printf("expected two levels\n"); return 4; }
        std::ofstream out(argv[3], std::ios::binary);
        out.write((const char*)vel0.data(), nv0 * 8);
        out.write((const char*)vel1.data(), nv1 * 8);
        for (int l = 0; l < 2; ++l) out.write((const char*)phi[l]->p, phi[l]->size() * 8);
        for (int l = 0; l < 2; ++l) out.write((const char*)gradphi[l]->p, gradphi[l]->size() * 8);
        double it = nodal_projector->stats().iters;
        out.write((const char*)&it, 8);
        std::printf("shim composite OK: %d iterations\n", nodal_projector->stats().iters);
    } catch (const std::runtime_error& e) {
        std::printf("amrex::Abort::%s\n", e.what());
        return 1;
    }
    return 0;
}

static int mac(int argc, char** argv)
{
    if (argc < 3) { std::printf("usage: shim_check mac n\n"); return 2; }
    abort_handler() = throwing_abort;
    const int N = std::atoi(argv[2]);
    const int n[3] = {N, N, N};
    Geometry g{{N, N, N}, {1.0 / N, 1.0 / N, 1.0 / N}, {true, true, false}};
    auto lo = get_mac_projection_bc(g.is_periodic, {BC::undefined, BC::undefined, BC::no_slip_wall});
    auto hi = get_mac_projection_bc(g.is_periodic, {BC::undefined, BC::undefined, BC::slip_wall});
    if (lo[2] != LinOpBCType::Neumann || hi[0] != LinOpBCType::Periodic) { std::printf("get_mac_projection_bc wrong\n"); return 3; }
    const size_t nu = (size_t)N * N * (N + 1), nc = (size_t)N * N * N;
    std::vector<double> u(nu), v(nu), w(nu), bx(nu), by(nu), bz(nu), phi(nc, 0.0);
    unsigned long long seed = 1234567ull;
    auto rnd = [&]() { seed ^= seed << 13; seed ^= seed >> 7; seed ^= seed << 17; return (double)(seed % 2000001ull) / 1.0e6 - 1.0; };
    for (auto& x : u) x = rnd();
    for (auto& x : v) x = rnd();
    for (auto& x : w) x = rnd();
    for (auto& x : bx) x = 0.75 + 0.25 * rnd();
    for (auto& x : by) x = 0.75 + 0.25 * rnd();
    for (auto& x : bz) x = 0.75 + 0.25 * rnd();
    // periodic x / y: one value per physical face; walls in z: no flow through them
    for (int k = 0; k < N; ++k) for (int j = 0; j < N; ++j) { u[((size_t)k * N + j) * (N + 1) + N] = u[((size_t)k * N + j) * (N + 1)]; bx[((size_t)k * N + j) * (N + 1) + N] = bx[((size_t)k * N + j) * (N + 1)]; }
    for (int k = 0; k < N; ++k) for (int i = 0; i < N; ++i) { v[((size_t)k * (N + 1) + N) * N + i] = v[((size_t)k * (N + 1)) * N + i]; by[((size_t)k * (N + 1) + N) * N + i] = by[((size_t)k * (N + 1)) * N + i]; }
    for (int j = 0; j < N; ++j) for (int i = 0; i < N; ++i) { w[(size_t)j * N + i] = 0.0; w[((size_t)N * N + j) * N + i] = 0.0; }
    const int flo[3] = {0, 0, 0};
    const int uhi[3] = {N, N - 1, N - 1}, vhi[3] = {N - 1, N, N - 1}, whi[3] = {N - 1, N - 1, N}, chi[3] = {N - 1, N - 1, N - 1};
    Fab fu(u.data(), flo, uhi, 1), fv(v.data(), flo, vhi, 1), fw(w.data(), flo, whi, 1), fphi(phi.data(), flo, chi, 1);
    Fab fbx(bx.data(), flo, uhi, 1), fby(by.data(), flo, vhi, 1), fbz(bz.data(), flo, whi, 1);
    (void)n;
    try {
        auto macproj = std::make_unique<MacProjector>(g);
        if (!macproj->needInitialization()) return 4;
        LPInfo lp_info;
        lp_info.setMaxCoarseningLevel(100);
        macproj->initProjector(lp_info, 0.01 / 1.0);                 // incflo.constant_density branch (:106-109)
        macproj->setDomainBC(lo, hi);
        std::vector<double> u0 = u, v0 = v, w0 = w;
        macproj->project(fu, fv, fw, 1e-11, 1e-14);
        const int it_const = macproj->stats().iters;
        u = u0; v = v0; w = w0;
        macproj->updateCoeffs({&fbx, &fby, &fbz});                   // variable density (:128)
        macproj->project(fphi, fu, fv, fw, 1e-11, 1e-14);            // m_use_mac_phi_in_godunov branch (:287-292)
        const int it_var = macproj->stats().iters;
        double dmax = 0.0, dsum = 0.0;
        for (int k = 0; k < N; ++k) for (int j = 0; j < N; ++j) for (int i = 0; i < N; ++i) {
            const double d = (u[((size_t)k * N + j) * (N + 1) + i + 1] - u[((size_t)k * N + j) * (N + 1) + i]) * N +
                             (v[((size_t)k * (N + 1) + j + 1) * N + i] - v[((size_t)k * (N + 1) + j) * N + i]) * N +
                             (w[((size_t)(k + 1) * N + j) * N + i] - w[((size_t)k * N + j) * N + i]) * N;
            dmax = std::max(dmax, std::fabs(d)); dsum += d;
        }
        const double bnorm = std::max(macproj->stats().rhsnorm, macproj->stats().resnorm0);
        if (!(dmax <= 2e-11 * bnorm + std::fabs(dsum) / nc)) { std::printf("divergence %g after the projection (bnorm %g)\n", dmax, bnorm); return 5; }
        std::vector<double> u1 = u0, v1 = v0, w1 = w0;
        Fab gu(u1.data(), flo, uhi, 1), gv(v1.data(), flo, vhi, 1), gw(w1.data(), flo, whi, 1);
        macproj->project(fphi, gu, gv, gw, 1e-10, 1e-14);            // warm start from the converged phi
        if (macproj->stats().iters != 0) { std::printf("warm start took %d iterations\n", macproj->stats().iters); return 6; }
        std::printf("shim mac OK: %d / %d V-cycles (constant / variable beta), max |div u| %.2e\n", it_const, it_var, dmax);
    } catch (const std::runtime_error& e) {
        std::printf("amrex::Abort::%s\n", e.what());
        return 1;
    }
    return 0;
}

static int eb(int argc, char** argv)
{
    if (argc < 16) { std::printf("usage: shim_check eb in.bin out.bin nx ny nz dx bclo(3) bchi(3) sigma ebflow\n"); return 2; }
    abort_handler() = throwing_abort;
    const int n[3] = {std::atoi(argv[4]), std::atoi(argv[5]), std::atoi(argv[6])};
    const double dx = std::atof(argv[7]);
    std::array<LinOpBCType, 3> lo, hi;
    Geometry g{{n[0], n[1], n[2]}, {dx, dx, dx}, {false, false, false}};
    for (int d = 0; d < 3; ++d) {
        lo[d] = (LinOpBCType)std::atoi(argv[8 + d]); hi[d] = (LinOpBCType)std::atoi(argv[11 + d]);
        g.is_periodic[d] = lo[d] == LinOpBCType::Periodic;
    }
    const double sigma = std::atof(argv[14]), ebflow = std::atof(argv[15]);
    const size_t nc = (size_t)n[0] * n[1] * n[2], ng1 = (size_t)(n[0] + 2) * (n[1] + 2) * (n[2] + 2);
    std::vector<double> vel(3 * ng1), vfrac(nc), intg(18 * nc), bnorm(3 * nc), bintg(8 * nc), ebvel(3 * ng1, -1.0);
    std::ifstream in(argv[2], std::ios::binary);
    in.read((char*)vel.data(), vel.size() * 8); in.read((char*)vfrac.data(), vfrac.size() * 8); in.read((char*)intg.data(), intg.size() * 8);
    in.read((char*)bnorm.data(), bnorm.size() * 8); in.read((char*)bintg.data(), bintg.size() * 8);
    if (!in) { std::printf("short input\n"); return 2; }
    try {
        EBFArrayBoxFactory factory{Fab::make(vfrac.data(), n, 0, 1), Fab::make(intg.data(), n, 0, 18), Fab::make(bnorm.data(), n, 0, 3),
                                   Fab::make(bintg.data(), n, 0, 8)};
        Fab fvel = Fab::make(vel.data(), n, 1, 3), feb = Fab::make(ebvel.data(), n, 1, 3);
        EBFlow flow;
        flow.enabled = ebflow != 0.0; flow.is_mag = true; flow.vel_mag = ebflow;
        if (flow.enabled) {   // set_eb_velocity(lev, time, *get_velocity_eb()[lev], 1)   (:130-136)
            IncfloEBNodalProjection inc(g, lo, hi, factory);
            inc.set_eb_flow(flow, 1, &feb, nullptr, nullptr);
        }
        LPInfo info;
        info.setMaxCoarseningLevel(100);
        auto nodal_projector = std::make_unique<EBNodalProjector>(fvel, sigma, g, factory, info);      // :187-188
        nodal_projector->setDomainBC(lo, hi);                                                            // :194
        if (flow.enabled) {
            // the linear operator takes eb_vel without ghost cells as well: pass the valid part through a ghost-free copy
            std::vector<double> ev(3 * nc);
            for (int c = 0; c < 3; ++c) for (int k = 0; k < n[2]; ++k) for (int j = 0; j < n[1]; ++j) for (int i = 0; i < n[0]; ++i)
                ev[((size_t)(c * n[2] + k) * n[1] + j) * n[0] + i] = ebvel[((size_t)(c * (n[2] + 2) + k + 1) * (n[1] + 2) + j + 1) * (n[0] + 2) + i + 1];
            Fab fev = Fab::make(ev.data(), n, 0, 3);
            nodal_projector->getLinOp().setEBInflowVelocity(0, fev);                                     // :196-201
            nodal_projector->project(1e-11, 1e-14);
        } else {
            nodal_projector->project(1e-11, 1e-14);                                                      // :215
        }
        auto phi = nodal_projector->getPhi();
        auto gradphi = nodal_projector->getGradPhi();
        const double iters = nodal_projector->stats().iters;
        std::ofstream out(argv[3], std::ios::binary);
        out.write((const char*)vel.data(), vel.size() * 8);
        out.write((const char*)phi[0]->p, phi[0]->size() * 8);
        out.write((const char*)gradphi[0]->p, gradphi[0]->size() * 8);
        out.write((const char*)ebvel.data(), ebvel.size() * 8);
        out.write((const char*)&iters, 8);
    } catch (const std::runtime_error& e) {
        std::printf("amrex::Abort::%s\n", e.what());
        return 1;
    }
    return 0;
}

int main(int argc, char** argv)
{
    if (argc >= 2 && !std::strcmp(argv[1], "eb")) return eb(argc, argv);
    if (argc >= 2 && !std::strcmp(argv[1], "mac")) return mac(argc, argv);
    if (argc >= 2 && !std::strcmp(argv[1], "host")) return host_checks();
    if (argc >= 2 && !std::strcmp(argv[1], "composite")) return composite(argc, argv);
    if (argc >= 2 && !std::strcmp(argv[1], "composite_mf")) return composite_mf(argc, argv);
    if (argc >= 2 && !std::strcmp(argv[1], "multibox")) return multibox(argc, argv);
    if (argc < 15 || std::strcmp(argv[1], "project")) { std::printf("usage: see header comment\n"); return 2; }
    abort_handler() = throwing_abort;
    const int n[3] = {std::atoi(argv[4]), std::atoi(argv[5]), std::atoi(argv[6])};
    const double dx = std::atof(argv[7]);
    std::array<LinOpBCType, 3> lo, hi;
    Geometry g;
    for (int d = 0; d < 3; ++d) {
        lo[d] = (LinOpBCType)std::atoi(argv[8 + d]); hi[d] = (LinOpBCType)std::atoi(argv[11 + d]);
        g.n_cell[d] = n[d]; g.dx[d] = dx; g.is_periodic[d] = lo[d] == LinOpBCType::Periodic;
    }
    const bool var = std::atoi(argv[14]) != 0;
    const size_t nv = (size_t)3 * (n[0] + 2) * (n[1] + 2) * (n[2] + 2), nc = (size_t)n[0] * n[1] * n[2];
    std::vector<double> vel(nv), sig(var ? nc : 0);
    std::ifstream in(argv[2], std::ios::binary);
    in.read((char*)vel.data(), nv * 8);
    if (var) in.read((char*)sig.data(), nc * 8);
    if (!in) { std::printf("short input\n"); return 3; }
    try {
        // the call sequence of incflo_apply_nodal_projection.cpp:181-219
        LPInfo info;
        info.setMaxCoarseningLevel(100);
        std::unique_ptr<NodalProjector> nodal_projector;
        std::vector<Fab> velv{Fab::make(vel.data(), n, 1, 3)};
        if (!var) nodal_projector = std::make_unique<NodalProjector>(velv, 0.37, std::vector<Geometry>{g}, info);
        else nodal_projector = std::make_unique<NodalProjector>(velv, std::vector<Fab>{Fab::make(sig.data(), n, 0, 1)}, std::vector<Geometry>{g}, info);
        nodal_projector->setDomainBC(lo, hi);
        nodal_projector->project(1e-11, 1e-14);
        auto phi = nodal_projector->getPhi();
        auto gradphi = nodal_projector->getGradPhi();
        std::ofstream out(argv[3], std::ios::binary);
        out.write((const char*)vel.data(), nv * 8);
        out.write((const char*)phi[0]->p, phi[0]->size() * 8);
        out.write((const char*)gradphi[0]->p, gradphi[0]->size() * 8);
        double it = nodal_projector->stats().iters;
        out.write((const char*)&it, 8);
        std::printf("shim project OK: %d V-cycles\n", nodal_projector->stats().iters);
    } catch (const std::runtime_error& e) {
        std::printf("amrex::Abort::%s\n", e.what());
        return 1;
    }
    return 0;
}
