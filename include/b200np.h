/*
 * b200np.h -- C ABI of the B200-native approximate nodal projection.
 *
 * Drop-in boundary for ONE hot path of AMReX-Fluids/incflo:
 *   incflo::ApplyNodalProjection            src/projection/incflo_apply_nodal_projection.cpp:29-267
 *     -> Hydro::NodalProjector (ctor, setDomainBC, project, getPhi, getGradPhi)   call sites :181-219
 *        -> amrex::MLMG / amrex::MLNodeLaplacian (un-vendored)
 *
 * Plain C: pointers + sizes only, no torch / AMReX types.  All field memory is
 * owned by the caller (AMReX arenas); every array comes with its own box
 * descriptor because ghost widths differ per MultiFab (src/incflo.H:714-724,
 * src/setup/incflo_arrays.cpp:9-26).  Array layout is amrex::Array4:
 *   idx = (i-lo.x) + nx*((j-lo.y) + ny*((k-lo.z) + nz*comp)),   nx = hi.x-lo.x+1 ...
 * Pointers may be device pointers (zero copy) or host pointers (staged H2D/D2H
 * inside the call); the kind is detected with cudaPointerGetAttributes.
 *
 * The solver runs ONLY as sm_100a CUDA kernels.  There is no CPU fallback: if
 * no CUDA device is usable every entry point returns B200NP_ERR_CUDA.
 */
#ifndef B200NP_H
#define B200NP_H

#ifdef __cplusplus
extern "C" {
#endif

#define B200NP_VERSION 1

/* amrex::LinOpBCType as produced by incflo::get_projection_bc
 * (src/projection/incflo_projection_bc.cpp:5-41):
 *   periodic -> PERIODIC; pi,po -> DIRICHLET; mi,dd,mixed -> INFLOW; sw,nsw -> NEUMANN */
enum b200np_bc { B200NP_BC_PERIODIC = 0, B200NP_BC_NEUMANN = 1, B200NP_BC_DIRICHLET = 2, B200NP_BC_INFLOW = 3 };

enum b200np_status {
    B200NP_OK = 0,
    B200NP_ERR_NOT_CONVERGED = 1, /* MLMG: failed to converge in maxiter (AMReX aborts)        */
    B200NP_ERR_DIVERGED = 2,      /* residual > 1e20 * max(rhsnorm,resnorm0) (AMReX aborts)    */
    B200NP_ERR_BAD_BC = 3,        /* get_projection_bc: undefined BC type (:36 aborts)         */
    B200NP_ERR_BAD_ARG = 4,
    B200NP_ERR_CUDA = 5,          /* no device / CUDA runtime error                            */
    B200NP_ERR_NCCL = 6,
    B200NP_ERR_UNSUPPORTED = 7,   /* general overset masks, AMR beyond one fine box at ratio 2, EB beyond one level / box */
    B200NP_ERR_PEER_TIMEOUT = 9,  /* slab path: a neighbour rank never raised its halo flag (it died or left the call early) */
    B200NP_ERR_INOUT_FLUX = 8     /* enforceInOutSolvability: inflow without outflow through the direction_dependent
                                     faces, or the reverse (AMReX-Hydro aborts)                 */
};

/* amrex::Geometry + domain BCs of one level (NodalProjector ctor + setDomainBC, :187-194) */
typedef struct {
    int    n_cell[3];  /* amr.n_cell: domain is cells [0,n_cell) (domain box lo = 0)   */
    double dx[3];      /* geom.CellSize()                                               */
    int    bc_lo[3];   /* enum b200np_bc                                                */
    int    bc_hi[3];
} b200np_geom;

/* nodal_proj.* keys.  Read by incflo: src/setup/init.cpp:172-177 (defaults
 * src/incflo.H:449-458); read by Hydro::NodalProjector::setOptions: documented
 * at src/incflo.H:436-445 and Docs/.../InputsMultigrid.rst:10-36. */
typedef struct {
    int    verbose;                 /* nodal_proj.verbose            0      */
    int    bottom_verbose;          /* nodal_proj.bottom_verbose     0      */
    int    maxiter;                 /* nodal_proj.maxiter            100    */
    int    bottom_maxiter;          /* nodal_proj.bottom_maxiter     100    */
    double bottom_rtol;             /* nodal_proj.bottom_rtol        1e-4   */
    double bottom_atol;             /* nodal_proj.bottom_atol        -1     */
    int    mg_max_coarsening_level; /* nodal_proj.mg_max_coarsening_level 100 */
    int    num_pre_smooth;          /* nodal_proj.num_pre_smooth     2      */
    int    num_post_smooth;         /* nodal_proj.num_post_smooth    2      */
    int    smooth_num_sweeps;       /* MLNodeLinOp m_smooth_num_sweeps 4    */
    int    bottom_solver;           /* 0 = bicgcg (default): BiCGStab, CG retry; 1 = smoother  */
    /* B200-specific knobs (not reference keys) */
    int    tile[3];                 /* smoother tile in nodes, default {64,16,64}: Gauss-Seidel
                                       inside a tile, previous-sweep values outside           */
    int    use_graph;               /* capture each V-cycle in a CUDA graph, default 1        */
} b200np_opts;

/* amrex::FArrayBox shape: the allocated (grown) box and component count */
typedef struct {
    int lo[3];
    int hi[3];
    int ncomp;
} b200np_fab;

/* what MLMG prints with nodal_proj.verbose >= 1, plus timers */
typedef struct {
    int    iters;            /* V-cycles                                    */
    int    nlevels;          /* MG levels                                   */
    int    bottom_iters;     /* BiCGStab iterations, summed                 */
    int    status;           /* enum b200np_status                          */
    double rhsnorm;          /* ||rhs||_inf                                 */
    double resnorm0;         /* initial ||rhs - L phi||_inf                 */
    double resnorm;          /* final                                       */
    double resnorm_hist[128];
    double ms_total;         /* device time of the whole call (CUDA events) */
    double ms_solve;         /* MLMG::solve part                            */
    double ms_h2d, ms_d2h;   /* staging when host pointers were passed      */
    long long h2d_bytes, d2h_bytes;
    long long launches;      /* kernels launched by this call               */
} b200np_stats;

typedef struct b200np b200np_t;

void b200np_default_opts(b200np_opts* o);

/* Hydro::NodalProjector::NodalProjector(vel, sigma, geom, LPInfo) + setOptions()
 * + setDomainBC(lo,hi)  (call sites :181-194).  The multigrid hierarchy is built
 * once and cached in the handle; incflo rebuilds it every projection (:181-193),
 * so re-using a handle across calls is a pure saving.
 * Single process / single GPU: device = CUDA ordinal. */
int b200np_create(b200np_t** out, const b200np_geom* geom, const b200np_opts* opts, int device);

/* Slab-decomposed multi-GPU variant (one process per GPU): this rank
 * owns cell planes [zlo, zhi) of the domain.  nccl_unique_id: the 128 bytes of an
 * ncclUniqueId created by rank 0 and broadcast by the host program. */
int b200np_create_dist(b200np_t** out, const b200np_geom* geom, const b200np_opts* opts, int device,
                       int rank, int nranks, const void* nccl_unique_id);

/* Helpers for the slab-decomposed variant.
 * b200np_nccl_unique_id: rank 0 creates the 128-byte ncclUniqueId the host program broadcasts
 *   (incflo would use MPI_Bcast / ParallelDescriptor::Bcast; bench.py uses torch.distributed).
 * b200np_slab_range: the z range of cells [cell_lo, cell_hi] and of uniquely owned node planes
 *   [node_lo, node_hi] of `rank` (the node plane shared with the upper neighbour belongs to the
 *   upper neighbour; the last rank of a non-periodic domain also owns the top boundary plane).
 *   Needs no GPU.
 * b200np_dist_plan: how many multigrid levels of `geom` stay slab-distributed over nranks ranks (the
 *   coarser ones are agglomerated by replication) and how many levels there are; min_planes <= 0 takes the
 *   default (64 cell planes per rank, B200NP_DIST_MIN_PLANES).  Replaces the agglomeration / consolidation
 *   decisions of MLLinOp::defineGrids.  Needs no GPU. */
int b200np_dist_plan(const b200np_geom* geom, int nranks, int min_planes, int max_coarsening_level, int* nlev_dist, int* nlev);
int b200np_nccl_unique_id(void* out128);
int b200np_slab_range(const b200np_geom* geom, int rank, int nranks, int* cell_lo, int* cell_hi, int* node_lo,
                      int* node_hi);

void b200np_destroy(b200np_t* h);

/* Run on the caller's CUDA stream (e.g. amrex::Gpu::gpuStream()) instead of the handle's own
 * non-blocking stream.  stream = a cudaStream_t cast to void*; NULL restores the own stream
 * (the legacy default stream cannot be captured into a graph, so it is never used). */
int b200np_set_stream(b200np_t* h, void* stream);

/* Hydro::NodalProjector::project(rtol, atol)   (:215) followed by getPhi() /
 * getGradPhi() (:218-219).
 *   vel   in/out: cell-centred, 3 comps, box grown by >= 1 ghost cell.  Valid
 *         cells are overwritten with u - sigma*grad(phi).  One ghost layer is an
 *         INPUT at non-periodic faces (0 at walls, inflow value at inflow faces,
 *         :137-163); periodic ghosts need not be filled (FillBoundary is internal).
 *   sigma cell-centred 1 comp (any ghost width; ghosts unused) or NULL => const_sigma
 *   phi   out, nodal box [0,n_cell] in every direction (may be NULL)
 *   gphi  out, cell-centred 3 comps, +grad(phi) (may be NULL)
 * In the slab-decomposed case every box is the rank's local box (cells
 * [zlo,zhi) in z, nodes [zlo,zhi] in z). */
int b200np_project(b200np_t* h, double* vel, const b200np_fab* vel_box, const double* sigma,
                   const b200np_fab* sigma_box, double const_sigma, double* phi, const b200np_fab* phi_box,
                   double* gphi, const b200np_fab* gphi_box, double rtol, double atol, b200np_stats* stats);

/* incflo::ApplyNodalProjection(density, time, scaling_factor, incremental)
 * (:29-93 and :95-267 fused; single level).
 *   velocity  in/out  ld.velocity  (3 comps, ng = nghost_state())
 *   velocity_o in     ld.velocity_o (used iff incremental || proj_for_small_dt, else may be NULL)
 *   density   in      density[lev] or NULL => incflo.constant_density with ro_0
 *   gp        in/out  ld.gp   (3 comps, 0 ghosts)
 *   p_nd      in/out  ld.p_nd (nodal)
 *   inflow_vel in     optional: same box as velocity; its first ghost layer at
 *                     INFLOW faces holds the IncfloVelFill values (src/prob/prob_bc.H:8-351)
 *                     to impose when set_inflow_bc = !proj_for_small_dt && !incremental (:81)
 * All velocity ghost cells are zeroed first, as vel.setBndry(0.0) does (:137). */
int b200np_apply_nodal_projection(b200np_t* h, double* velocity, const b200np_fab* vel_box,
                                  const double* velocity_o, const double* density, const b200np_fab* rho_box,
                                  double ro_0, double* gp, const b200np_fab* gp_box, double* p_nd,
                                  const b200np_fab* p_box, const double* inflow_vel, double scaling_factor,
                                  int incremental, int proj_for_small_dt, double rtol, double atol,
                                  b200np_stats* stats);

/* ---- multi-box MultiFabs (amr.max_grid_size < domain: every reference deck, e.g. test_no_eb_3d/benchmark.rayleigh_taylor:16) ----
 * amrex::MultiFab as this rank sees it: the FArrayBoxes of its local MFIter.  box[f] is the ALLOCATED box of fab f
 * (fab.box(): the valid box grown by ngrow, in the MultiFab's own index space -- cell-centred, or nodal for p_nd / phi),
 * data[f] its dataPtr().  All pointers of one MultiFab are of the same kind (device, or host => staged inside the call).
 * The valid boxes must tile this rank's part of the domain (the whole domain on one GPU, the rank's z slab otherwise) --
 * what an amrex::BoxArray guarantees.  The solver gathers the boxes into its slab arrays with one fused pass per
 * field and scatters the results back the same way; results are bit-identical to the single-box calls. */
typedef struct {
    int nfabs;                 /* MultiFab::local_size()                                  */
    int ngrow;                 /* MultiFab::nGrow()                                       */
    int ncomp;                 /* MultiFab::nComp()                                       */
    const b200np_fab* box;     /* [nfabs] allocated box of each fab (ncomp field ignored)  */
    double* const* data;       /* [nfabs] fab.dataPtr()                                   */
} b200np_mfab;
/* b200np_project over MultiFabs.  After the call every cell of vel's fabs that lies inside the domain grown by one cell
 * holds the projected velocity / the BC ghost value (interior ghost cells as after FillBoundary), sigma may be NULL. */
int b200np_project_mf(b200np_t* h, const b200np_mfab* vel, const b200np_mfab* sigma, double const_sigma, const b200np_mfab* phi,
                      const b200np_mfab* gphi, double rtol, double atol, b200np_stats* stats);
/* b200np_apply_nodal_projection over incflo::LevelData's MultiFabs (velocity, velocity_o, density, gp, p_nd).  On
 * return the ghost cells of velocity that lie inside the domain hold the neighbour boxes' projected values (as after
 * FillBoundary; not across periodic faces beyond the first layer), the first layer outside the domain the BC value (0,
 * or the inflow fill) and everything further out 0 (vel.setBndry(0.0), :137). */
int b200np_apply_nodal_projection_mf(b200np_t* h, const b200np_mfab* velocity, const b200np_mfab* velocity_o,
                                     const b200np_mfab* density, double ro_0, const b200np_mfab* gp, const b200np_mfab* p_nd,
                                     const b200np_mfab* inflow_vel, double scaling_factor, int incremental, int proj_for_small_dt,
                                     double rtol, double atol, b200np_stats* stats);

/* ---- composite (two AMR level) projection: BASELINE configs[3], amr.max_level = 1 -------------------
 * incflo::ApplyNodalProjection loops over lev = 0..finest_level (:101-121, :130-164, :221-266) and hands
 * Hydro::NodalProjector the vectors vel[], sigma[], Geom(0,finest_level) (:181-192); AMReX's MLMG then
 * solves the composite problem (MLMG::oneIter multi-level branch, MLNodeLaplacian::reflux / compRHS /
 * interpolationAmr; un-vendored, restated in oracle/composite.py).  Supported here: ONE fine box at
 * amr.ref_ratio = 2 (every deck).  Per direction and side the box either ends at least one coarse cell
 * inside the domain (a coarse/fine interface), or spans a periodic direction completely (e.g. a refined
 * slab around the rayleigh_taylor interface), or touches a wall (Neumann) / outflow (Dirichlet) face, whose
 * BC the fine level then inherits.  A box on an inflow face, or touching the periodic seam without
 * spanning the direction, several boxes, or a third level
 * returns B200NP_ERR_UNSUPPORTED.  fine_lo / fine_hi: the covered COARSE cells (inclusive).  Boxes of
 * level-1 arrays are in FINE index space (fine cells 2*fine_lo .. 2*fine_hi+1), as amrex::MultiFab
 * boxes of level 1 are.  Coarse cells under the fine box: their input velocity never counts; on return
 * velocity, gp (gphi) there are the average of the fine values (amrex::average_down, :258-266 and
 * NodalProjector::averageDown) and phi / p_nd on the covered coarse nodes is the fine value. */
typedef struct b200np_composite b200np_composite_t;
int  b200np_composite_create(b200np_composite_t** out, const b200np_geom* geom0, const int fine_lo[3], const int fine_hi[3],
                             const b200np_opts* opts, int device);
void b200np_composite_destroy(b200np_composite_t* c);
int  b200np_composite_set_stream(b200np_composite_t* c, void* stream);
/* the single-level handle of AMR level 0 / 1 (test hooks b200np_level_* / b200np_time_op; owned by c) */
b200np_t* b200np_composite_level(b200np_composite_t* c, int amr_level);
/* Hydro::NodalProjector::project over two levels + getPhi / getGradPhi (:215-219).  sigma0 and sigma1 are
 * both given or both NULL (=> const_sigma).  One ghost layer of vel0 is an input at non-periodic domain
 * faces as in b200np_project; the ghost cells of vel1 are zeroed (vel.setBndry(0.0), :137). */
int  b200np_composite_project(b200np_composite_t* c, double* vel0, const b200np_fab* vel0_box, double* vel1,
                              const b200np_fab* vel1_box, const double* sigma0, const b200np_fab* sigma0_box,
                              const double* sigma1, const b200np_fab* sigma1_box, double const_sigma, double* phi0,
                              const b200np_fab* phi0_box, double* phi1, const b200np_fab* phi1_box, double* gphi0,
                              const b200np_fab* gphi0_box, double* gphi1, const b200np_fab* gphi1_box, double rtol, double atol,
                              b200np_stats* stats);
/* incflo::ApplyNodalProjection with finest_level = 1: every per-level argument of
 * b200np_apply_nodal_projection becomes an array indexed by AMR level. */
int  b200np_composite_apply_nodal_projection(b200np_composite_t* c, double* const velocity[2], const b200np_fab* const vel_box[2],
                                             const double* const velocity_o[2], const double* const density[2],
                                             const b200np_fab* const rho_box[2], double ro_0, double* const gp[2],
                                             const b200np_fab* const gp_box[2], double* const p_nd[2],
                                             const b200np_fab* const p_box[2], const double* inflow_vel0, double scaling_factor,
                                             int incremental, int proj_for_small_dt, double rtol, double atol, b200np_stats* stats);

/* The same over multi-box MultiFabs (every deck chops both levels with amr.max_grid_size; e.g.
 * test_no_eb_3d/benchmark.bouss_bubble_god:20): velocity[l], gp[l], p_nd[l], density[l], velocity_o[l] are the MultiFabs of
 * incflo::LevelData on AMR level l as in b200np_apply_nodal_projection_mf.  The valid boxes of level 0 tile the domain, those of
 * level 1 tile the fine box [2*fine_lo, 2*fine_hi+1] (fine index space) -- a refined region that is a union of several
 * rectangles is still B200NP_ERR_UNSUPPORTED.  Results are bit-identical to the single-box call. */
int  b200np_composite_apply_nodal_projection_mf(b200np_composite_t* c, const b200np_mfab* const velocity[2],
                                                const b200np_mfab* const velocity_o[2], const b200np_mfab* const density[2],
                                                double ro_0, const b200np_mfab* const gp[2], const b200np_mfab* const p_nd[2],
                                                const b200np_mfab* inflow_vel0, double scaling_factor, int incremental,
                                                int proj_for_small_dt, double rtol, double atol, b200np_stats* stats);

/* ---- MAC projection: Hydro::MacProjector over amrex::MLMG / MLABecLaplacian --------------------------------------
 * Call sites: src/convection/incflo_compute_MAC_projected_velocities.cpp:69-129 (inv_rho on faces = dt / rho,
 * macproj->initProjector(lp_info, inv_rho) | initProjector(ba, dm, lp_info, dt / ro_0), setDomainBC(get_mac_projection_bc),
 * updateCoeffs / updateBeta) and :280-299 (project(mac_phi, rtol, atol) | project(rtol, atol)); keys mac_proj.mg_rtol,
 * mg_atol, mg_max_coarsening_level (src/setup/init.cpp:165-170).  One AMR level, one box, one GPU.
 * geom.bc_lo / bc_hi: the LinOpBCType of incflo::get_mac_projection_bc (src/projection/incflo_projection_bc.cpp:43-79):
 * PERIODIC, NEUMANN (walls, mass inflow, direction_dependent; B200NP_BC_INFLOW is accepted and means NEUMANN), DIRICHLET
 * (pressure in/outflow).  Robin (mixed faces) returns B200NP_ERR_UNSUPPORTED from the C++ mirror.
 * opts: maxiter, bottom_maxiter (MLMG defaults 200), bottom_rtol / bottom_atol, mg_max_coarsening_level, num_pre_smooth,
 * num_post_smooth, verbose are used; opts == NULL takes these MLMG defaults. */
typedef struct b200mac b200mac_t;
int  b200mac_create(b200mac_t** out, const b200np_geom* geom, const b200np_opts* opts, int device);
void b200mac_destroy(b200mac_t* h);
int  b200mac_nlevels(const b200mac_t* h);
int  b200mac_set_stream(b200mac_t* h, void* stream);   /* as b200np_set_stream */
/* initProjector / updateCoeffs (bx, by, bz: face-centred dt / rho, boxes in face index space: x faces [0,nx] x [0,ny) x
 * [0,nz) ...) or, with all three NULL, initProjector(..., const_beta) / updateBeta(const_beta). */
int  b200mac_set_coeffs(b200mac_t* h, const double* bx, const b200np_fab* bx_box, const double* by, const b200np_fab* by_box,
                        const double* bz, const b200np_fab* bz_box, double const_beta);
/* project: rhs = -div(u_mac); MLMG solve of -div(b grad phi) = rhs to max(atol, rtol * max(|rhs|, |res0|)); u_mac -= b grad phi
 * on every face (boundary faces with the BC ghost cell: nothing at Neumann faces).  mac_phi (optional cell array): the
 * initial guess when phi_is_initial_guess (project(mac_phi, ...), :287-292, incflo passes zeros), and phi on return. */
int  b200mac_project(b200mac_t* h, double* umac, const b200np_fab* u_box, double* vmac, const b200np_fab* v_box, double* wmac,
                     const b200np_fab* w_box, double* mac_phi, const b200np_fab* phi_box, int phi_is_initial_guess, double rtol,
                     double atol, b200np_stats* stats);
/* The same two calls over multi-box MultiFabs (b200np_mfab below; every reference deck runs with amr.max_grid_size = 16): bx / by / bz and
 * umac / vmac / wmac face-centred (the valid box of a fab = its cells' box + the far face of its direction, shared with the neighbour box),
 * mac_phi cell-centred.  The fabs are gathered into one array per field, projected, and the valid faces / cells scattered back:
 * bit-identical to the single-box call. */
int  b200mac_set_coeffs_mf(b200mac_t* h, const b200np_mfab* bx, const b200np_mfab* by, const b200np_mfab* bz);
int  b200mac_project_mf(b200mac_t* h, const b200np_mfab* umac, const b200np_mfab* vmac, const b200np_mfab* wmac,
                        const b200np_mfab* mac_phi, int phi_is_initial_guess, double rtol, double atol, b200np_stats* stats);
/* test hooks: op 0 smooth (arg MLMG smooth calls: cor = in_a, res = in_b), 1 residual (in_b - A in_a), 2 restriction of in_a to
 * level lev + 1, 3 in_a + interpolation of in_b (level lev + 1), 4 bottom solve of in_b on the coarsest level.  Host arrays,
 * dense cell layout (nz, ny, nx). */
int  b200mac_level_op(b200mac_t* h, int lev, int op, int arg, const double* in_a, const double* in_b, double* out);
int  b200mac_level_dims(const b200mac_t* h, int lev, int n_cell[3]);

/* ---- EB (cut cell) nodal projection: Hydro::NodalProjector over MLMG / MLNodeLaplacian built with an EBFArrayBoxFactory ------
 * What incflo runs under AMREX_USE_EB (BASELINE configs[4], test_3d/benchmark.channel_cylinder-x): call sites
 * src/projection/incflo_apply_nodal_projection.cpp:130-136 (set_eb_velocity / density / tracer), :181-194 (projector, whose ctor takes
 * the EB factory from vel[lev]->Factory()), :196-201 (getLinOp().setEBInflowVelocity), :215-266 (project, copy-out).
 * One AMR level, one box, one GPU.  geom.dx must be isotropic (AMReX's EB support asserts dx == dy == dz).
 * The operator is the Q1 stiffness matrix integrated over the fluid part of every cell; the multigrid is AMReX's "RAP" strategy
 * (Galerkin coarse stencils); see csrc/b200eb.cu and oracle/eb_oracle.py. */
typedef struct b200eb b200eb_t;
int  b200eb_create(b200eb_t** out, const b200np_geom* geom, const b200np_opts* opts, int device);
void b200eb_destroy(b200eb_t* h);
int  b200eb_nlevels(const b200eb_t* h);
int  b200eb_set_stream(b200eb_t* h, void* stream);     /* as b200np_set_stream */
/* The EB data the linear operator takes from the factory, per cell of the level (cell-centred boxes that cover the domain):
 *   vfrac  EBFArrayBoxFactory::getVolFrac()            1 comp
 *   intg   MLNodeLaplacian::m_integral (buildIntegral) 18 comps, integrals of x y z x2 y2 z2 xy xz yz x2y x2z xy2 y2z xz2 yz2 x2y2
 *          x2z2 y2z2 over the fluid part of the cell, cell-local coordinates in [-1/2, 1/2] (amrex i_S_* order)
 * Uncut cells: vfrac = 1, int x2 = 1/12, int x2y2 = 1/144, odd ones 0; covered cells: all 0. */
int  b200eb_set_geometry(b200eb_t* h, const double* vfrac, const b200np_fab* vfrac_box, const double* intg, const b200np_fab* intg_box);
/* nodal_projector->getLinOp().setEBInflowVelocity(lev, eb_vel) (:196-201): adds dxinv * (u_eb . n) * int_{EB face} N_a dA to the rhs.
 *   eb_vel 3 comps (get_velocity_eb), bnorm 3 comps (getBndryNormal, pointing out of the fluid), bintg 8 comps: integrals of
 *   1 x y z xy xz yz xyz over the EB face inside the cell (comp 0 = getBndryArea; amrex i_B_* order).  eb_vel == NULL clears it. */
int  b200eb_set_eb_inflow_velocity(b200eb_t* h, const double* eb_vel, const b200np_fab* vel_box, const double* bnorm,
                                   const b200np_fab* bnorm_box, const double* bintg, const b200np_fab* bintg_box);
/* incflo::set_eb_velocity / set_eb_density / set_eb_tracer (src/boundary_conditions/incflo_set_bcs.cpp:195-285, :287-358, :360-431):
 * the eb_flow.* inputs; every output array is zeroed, cut cells (EBCellFlag::isSingleValued, 0 < vfrac < 1) get the value -- for
 * the velocity either eb_flow.velocity or -normal * eb_flow.vel_mag -- masked by the direction test
 * -1 - (normal_tol + eps_float) <= bnorm . eb_flow.normal <= -1 + (normal_tol + eps_float), and nghost layers are filled across periodic
 * faces (FillBoundary).  Any of the three outputs may be NULL. */
typedef struct {
    int    has_normal;     /* eb_flow.normal given        */
    double normal[3];
    double normal_tol;     /* eb_flow.normal_tol          */
    int    is_mag;         /* velocity given as magnitude */
    double vel_mag;
    double velocity[3];
    double density;
    int    ntrac;          /* <= 8                        */
    double tracer[8];
} b200eb_flow;
int  b200eb_set_eb_flow(b200eb_t* h, const b200eb_flow* f, int nghost, const double* bnorm, const b200np_fab* bnorm_box, double* eb_vel,
                        const b200np_fab* vel_box, double* eb_density, const b200np_fab* density_box, double* eb_tracer,
                        const b200np_fab* tracer_box);
/* NodalProjector::project(rtol, atol) + getPhi() / getGradPhi(): same arguments as b200np_project.  vel: valid cells become
 * u - sigma * (cell average of grad phi over the fluid), 0 in covered cells; one ghost layer is an input at non-periodic faces.
 * phi: nodal box [0, n_cell] (0 on nodes inside the body), gphi: cell average of grad phi over the fluid part. */
int  b200eb_project(b200eb_t* h, double* vel, const b200np_fab* vel_box, const double* sigma, const b200np_fab* sigma_box,
                    double const_sigma, double* phi, const b200np_fab* phi_box, double* gphi, const b200np_fab* gphi_box, double rtol,
                    double atol, b200np_stats* stats);
/* incflo::ApplyNodalProjection under AMREX_USE_EB: same arguments and semantics as b200np_apply_nodal_projection */
int  b200eb_apply_nodal_projection(b200eb_t* h, double* velocity, const b200np_fab* vel_box, const double* velocity_o,
                                   const double* density, const b200np_fab* rho_box, double ro_0, double* gp, const b200np_fab* gp_box,
                                   double* p_nd, const b200np_fab* p_box, const double* inflow_vel, double scaling_factor,
                                   int incremental, int proj_for_small_dt, double rtol, double atol, b200np_stats* stats);
/* The same over multi-box MultiFabs (b200np_mfab; amr.max_grid_size < domain, e.g. test_3d/benchmark.channel_sphere: 16): the fabs are gathered
 * into one array per field, the single-box path runs, the results are scattered back -- bit-identical to the single-box calls.  After the
 * call every cell of the velocity fabs inside the domain grown by one cell holds the new velocity / the BC ghost value, the cells beyond 0. */
int  b200eb_set_geometry_mf(b200eb_t* h, const b200np_mfab* vfrac, const b200np_mfab* intg);
int  b200eb_project_mf(b200eb_t* h, const b200np_mfab* vel, const b200np_mfab* sigma, double const_sigma, const b200np_mfab* phi,
                       const b200np_mfab* gphi, double rtol, double atol, b200np_stats* stats);
int  b200eb_apply_nodal_projection_mf(b200eb_t* h, const b200np_mfab* velocity, const b200np_mfab* velocity_o, const b200np_mfab* density,
                                      double ro_0, const b200np_mfab* gp, const b200np_mfab* p_nd, const b200np_mfab* inflow_vel,
                                      double scaling_factor, int incremental, int proj_for_small_dt, double rtol, double atol,
                                      b200np_stats* stats);
/* test hooks (host arrays, natural node order (nnz, nny, nnx), nn = n_cell in a periodic direction, n_cell + 1 otherwise):
 * build_stencils: MLNodeLaplacian::buildStencil for a sigma without projecting; level_stencil: the 13 forward entries
 * (offset t = (di+1) + 3(dj+1) + 9(dk+1), t = 14..26) + the diagonal of a level, (14, nnz, nny, nnx);
 * level_op: 0 smooth (arg MLMG smooth calls: x = in_a, rhs = in_b), 1 residual in_b - A in_a, 2 restriction of in_a to level lev + 1,
 * 3 in_a + interpolation of in_b (level lev + 1), 4 bottom solve of in_b on the coarsest level, 5 A in_a; compute_rhs: D u (+ EB inflow). */
int  b200eb_build_stencils(b200eb_t* h, const double* sigma, const b200np_fab* sigma_box, double const_sigma);
int  b200eb_level_stencil(b200eb_t* h, int lev, double* out);
int  b200eb_level_op(b200eb_t* h, int lev, int op, int arg, const double* in_a, const double* in_b, double* out);
int  b200eb_level_dims(const b200eb_t* h, int lev, int n_cell[3], int n_node[3]);
int  b200eb_compute_rhs(b200eb_t* h, const double* vel, const b200np_fab* vel_box, double* out);
/* measurement hook: reps x { arg smooth calls (op 0) | one residual (op 1) } on level lev, CUDA events; ms per repetition */
int  b200eb_time_op(b200eb_t* h, int lev, int op, int arg, int reps, double* ms);

/* IncfloVelFill (src/prob/prob_bc.H:8-351) evaluated by the library: after this call,
 * b200np_apply_nodal_projection with inflow_vel == NULL fills the first ghost layer of the velocity at INFLOW
 * faces itself (:138-163, PhysBCFunct<GpuBndryFuncFab<IncfloVelFill>> with nghost = 1) from
 *   bcv_vel[6][3]  m_bc_velocity: boundary velocity per amrex::Orientation (x-lo, y-lo, z-lo, x-hi, y-hi, z-hi),
 *   probtype       the profile of the normal component (16, 31, 311, 32, 322, 33, 333, 41, 42, 43; else bcv_vel),
 *   time           probtype 42.
 * Face kinds come from b200np_set_face_types (default: mass inflow); the special-case blocks of probtypes 1101 / 1102
 * (prob_bc.H:86-92, :140-146, :243-251) are evaluated as written there.  bcv_vel == NULL switches the profile off again. */
int b200np_set_inflow_profile(b200np_t* h, int probtype, const double* bcv_vel, double time);

/* incflo's own face types where they matter beyond the LinOpBCType of get_projection_bc (BC enum, src/incflo.H:662-665;
 * parsed in src/boundary_conditions/boundary_conditions.cpp:20-135).  face_type[6] in amrex::Orientation order
 * (x-lo, y-lo, z-lo, x-hi, y-hi, z-hi); only B200NP_BC_INFLOW faces may carry a non-default type.
 *   DEFAULT              mass inflow ("mi"): IncfloVelFill imposes the boundary velocity (BCType::ext_dir)
 *   DIRECTION_DEPENDENT  "dd": IncfloVelFill imposes the boundary velocity where the profile points into the domain and
 *                        copies the first interior cell where it points out (prob_bc.H:93-109 ...), and
 *                        HydroUtils::enforceInOutSolvability rescales the outflow so that the net flux through these faces
 *                        vanishes (incflo_apply_nodal_projection.cpp:166-179; has_inout_bndry).  Returns
 *                        B200NP_ERR_INOUT_FLUX from the projection where AMReX-Hydro aborts.
 *   MIXED                "mixed" (probtypes 1100/1101/1102): inflow on one half of the face, outflow on the other.  The
 *                        outflow half is imposed as Dirichlet nodes through the solver's overset mask
 *                        (incflo::make_nodalBC_mask, src/boundary_conditions/incflo_set_bcs.cpp:10-53 with
 *                        prob_set_BC_MF, src/prob/prob_bc.cpp:9-101): on a low-side face the nodes with
 *                        idx[mixed_split_dir] <= mixed_half_num_cells, on a high-side face those with idx > half;
 *                        mixed_half_num_cells = domain.length(mixed_split_dir) / 2 in incflo.
 * Replaces nodal_projector->getLinOp().setOversetMask(lev, make_nodalBC_mask(lev)) (:204-213) for the masks incflo can
 * produce; b200np_check_overset_mask verifies a caller-built mask against it. */
enum b200np_face_type { B200NP_FACE_DEFAULT = 0, B200NP_FACE_DIRECTION_DEPENDENT = 1, B200NP_FACE_MIXED = 2 };
int b200np_set_face_types(b200np_t* h, const int face_type[6], int mixed_split_dir, int mixed_half_num_cells);
/* MLNodeLaplacian::setOversetMask(lev, mask): mask is the caller's nodal int array (1 = solve, 0 = known / Dirichlet) on
 * mask_box (nodal box covering this rank's nodes).  Supported masks are exactly those b200np_set_face_types describes;
 * returns B200NP_OK when `mask` equals the mask in effect, B200NP_ERR_UNSUPPORTED otherwise. */
int b200np_check_overset_mask(b200np_t* h, const int* mask, const b200np_fab* mask_box);
/* influx / outflux (sum |u_n| dS) found by the last projection's enforceInOutSolvability */
int b200np_inout_flux(const b200np_t* h, double* influx, double* outflux);

const char* b200np_strerror(int status);
int b200np_version(void);

/* ---- test / profiling hooks: run one multigrid building block on device-resident
 * level arrays so that each kernel can be compared with the oracle.  Host arrays
 * are in the unique-node layout (nn = n in periodic directions, n+1 otherwise;
 * i fastest; no ghosts) or plain cell layout. ---- */
enum b200np_array { B200NP_A_SOL = 0, B200NP_A_RHS = 1, B200NP_A_RES = 2, B200NP_A_COR = 3, B200NP_A_RESCOR = 4,
                    B200NP_A_SIGMA = 5 };
enum b200np_op { B200NP_OP_SMOOTH = 0,   /* cor <- nsweeps sweeps on (cor, res)         */
                 B200NP_OP_RESIDUAL = 1, /* rescor <- res - L cor                       */
                 B200NP_OP_RESTRICT = 2, /* res[lev+1] <- R rescor[lev]                 */
                 B200NP_OP_INTERP = 3,   /* cor[lev] += P cor[lev+1]                    */
                 B200NP_OP_BOTTOM = 4,   /* cor[bottom] <- bottom solve of res[bottom]  */
                 B200NP_OP_VCYCLE = 5,   /* one V-cycle on (cor, res) from level 0      */
                 B200NP_OP_COARSEN_SIGMA = 6 };
/* B200NP_OP_SMOOTH: arg = number of sweeps, optionally | B200NP_SMOOTH_ZERO_START: "cor is zero" smooth call as in the
 * V-cycle's pre-smooth (MLMG::mgVcycle sets cor = 0 first, A.9) -- the first sweep does not read cor at all */
#define B200NP_SMOOTH_ZERO_START 0x10000
int b200np_nlevels(const b200np_t* h);
/* how the slab halos travel: 0 = single GPU (none), 1 = NVLink peer memory (CUDA IPC mapped neighbour arenas, stores /
 * loads issued by the solver kernels), 2 = grouped ncclSend/ncclRecv (fallback when a rank cannot map its neighbours) */
int b200np_halo_transport(const b200np_t* h);
/* how the neighbours' memory was mapped when the transport is 1: 1 = cuMemCreate allocation shared as a POSIX file
 * descriptor over a UNIX socket (the default; what NCCL itself does), 2 = legacy cudaIpc handles; 0 = not mapped */
int b200np_peer_map(const b200np_t* h);
int b200np_level_dims(const b200np_t* h, int lev, int n_cell[3], int n_node[3]);
int b200np_set_sigma(b200np_t* h, const double* sigma, const b200np_fab* sigma_box, double const_sigma);
int b200np_level_set(b200np_t* h, int lev, int which, const double* host);
int b200np_level_get(b200np_t* h, int lev, int which, double* host);
int b200np_level_op(b200np_t* h, int lev, int op, int arg);
/* time `reps` back-to-back launches of one op with CUDA events; returns ms per launch in *ms */
int b200np_time_op(b200np_t* h, int lev, int op, int arg, int reps, double* ms);

#ifdef __cplusplus
}
#endif
#endif
