/*
 * nodal_oracle.h -- CPU oracle for incflo's approximate nodal projection.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (incflo_b200/, the
 * C-ABI library libb200np.so) may include, link or call this.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker / CPU baseline.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in AMReX
 * (Src/LinearSolvers/MLMG: MLMG, MLNodeLaplacian, MLNodeLinOp, MLCGSolver) and
 * AMReX-Hydro (Projections/hydro_NodalProjector), neither of which is vendored
 * in /root/reference nor pinned to a version (README.md:14-28,
 * CMakeLists.txt:85-180, .github/workflows/gcc.yml:25-34).  The reference ships
 * no golden vectors for the projection.  This file is a
 * restatement of the published algorithm anchored on the
 * reference's own call site, src/projection/incflo_apply_nodal_projection.cpp,
 * and pinned only by independent mathematical checks (tests/test_oracle_*.py):
 * operator == assembled Q1 finite-element stiffness, D = -G^T adjointness,
 * direct sparse solve == multigrid solve, 2nd-order convergence on the
 * reference's Taylor-Green initial condition.
 */
#ifndef NODAL_ORACLE_H
#define NODAL_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* LinOpBCType values used by incflo::get_projection_bc
 * (src/projection/incflo_projection_bc.cpp:5-41) */
enum { ORC_BC_PERIODIC = 0, ORC_BC_NEUMANN = 1, ORC_BC_DIRICHLET = 2, ORC_BC_INFLOW = 3 };

/* smoother orderings */
enum {
    ORC_SM_LEX      = 0, /* lexicographic Gauss-Seidel, one box (AMReX CPU path)   */
    ORC_SM_COLOR8   = 1, /* 8-colour Gauss-Seidel c=(i&1)+2(j&1)+4(k&1) (AMReX GPU path, A.4)   */
    ORC_SM_COLOR4XY = 2, /* 4 colours c=(i&1)+2(j&1); same-colour nodes couple only along z     */
    ORC_SM_JACOBI   = 3, /* weighted Jacobi, omega = 2/3 (AMReX use_gauss_seidel=0)             */
    ORC_SM_PLANE4   = 5, /* planes in ascending k; inside a plane the 4 colours c=(i&1)+2(j&1), 0..3   */
    ORC_SM_BOX      = 4  /* box-decomposed: GS inside boxes of `box` cells, stale values outside
                            (AMReX multi-box semantics: no halo refresh between the sweeps of
                            one smooth call when box_stale_per_call=1) */
};

typedef struct {
    int    n[3];            /* cells per direction                                    */
    double dx[3];
    int    bclo[3], bchi[3];
    /* nodal_proj.* (src/setup/init.cpp:172-177, src/incflo.H:449-458 and the
     * NodalProjector::setOptions keys documented at src/incflo.H:436-445) */
    int    max_coarsening_level;   /* nodal_proj.mg_max_coarsening_level, default 100 */
    int    maxiter;                /* nodal_proj.maxiter, default 100                 */
    int    bottom_maxiter;         /* default 100                                     */
    double bottom_rtol;            /* default 1e-4                                    */
    double bottom_atol;            /* default -1                                      */
    int    nu1, nu2;               /* pre/post smooth calls, default 2/2              */
    int    nsweeps;                /* sweeps per smooth call, default 4 (A.4)         */
    int    smoother;               /* ORC_SM_*                                        */
    int    box[3];                 /* ORC_SM_BOX: box size in nodes                   */
    int    box_order;              /* in-box ordering: LEX / COLOR8 / COLOR4XY / PLANE4 */
    int    box_stale_per_call;     /* 1: snapshot once per smooth call; 0: per sweep  */
    int    verbose;
    int    box_amrex;              /* ORC_SM_BOX only.  1 = AMReX grid semantics (the reference's CPU path):
                                      box[] is amr.max_grid_size in CELLS on level 0 and is coarsened together with the
                                      multigrid level (never below 2 cells), the top node plane of a non-periodic
                                      direction belongs to the last box, and the GPU smoother's z-chunk rule is NOT
                                      applied.  Use with box_order = ORC_SM_LEX, box_stale_per_call = 1.            */
    /* incflo BC::mixed faces (src/boundary_conditions/incflo_set_bcs.cpp:10-53 make_nodalBC_mask +
     * src/prob/prob_bc.cpp:9-101 prob_set_BC_MF, probtypes 1100/1101/1102): the projection sees LinOpBCType::inflow on
     * the face (incflo_projection_bc.cpp:23-27) plus an overset mask whose zeros are Dirichlet nodes -- on a low-side
     * face the nodes with idx[mix_dir] <= mix_half, on a high-side face those with idx[mix_dir] > mix_half
     * (mix_half = domain.length(mix_dir) / 2).  Coarser multigrid levels take the mask by injection (node 2i). */
    int    mixed_lo[3], mixed_hi[3];
    int    mix_dir, mix_half;
} orc_params;

typedef struct {
    int    iters;          /* V-cycles used                        */
    int    nlevels;        /* multigrid levels                     */
    int    bottom_iters;   /* total BiCGStab iterations            */
    int    status;         /* 0 ok, 1 not converged, 2 diverged    */
    double rhsnorm;        /* ||rhs||_inf after solvability fix    */
    double resnorm0;       /* initial residual                     */
    double resnorm;        /* final residual ||rhs - L phi||_inf   */
    double resnorm_hist[128];
    double t_solve;        /* seconds in the MLMG solve            */
    double t_total;        /* seconds in project()                 */
} orc_stats;

void orc_default_params(orc_params* p);
/* OpenMP threads used by every orc_* call from now on (torchrun exports OMP_NUM_THREADS=1); returns the
 * number actually in effect (omp_get_max_threads) */
int  orc_set_num_threads(int n);

/*
 * Hydro::NodalProjector::project.  Arrays are Fortran order
 * (i fastest), component outermost, as AMReX Array4.
 *   vel    : cell-centred, 3 comps, box grown by ng>=1 ghost cells; in/out.
 *            One ghost layer is read: caller provides wall (0) / inflow
 *            values at non-periodic faces; periodic ghosts are filled here.
 *   sigma  : cell-centred n[0]*n[1]*n[2], no ghosts, or NULL => const_sigma.
 *   phi    : nodal (n+1)^3, out.
 *   gphi   : cell-centred 3 comps, no ghosts, out (+grad phi).
 *   rhs_out: optional nodal (n+1)^3 copy of the rhs D u (after x2 scaling,
 *            before the solvability offset), may be NULL.
 */
int orc_project(const orc_params* p, double* vel, int ng, const double* sigma, double const_sigma,
                double* phi, double* gphi, double* rhs_out, double rtol, double atol, orc_stats* st);

/*
 * incflo::ApplyNodalProjection, 4-arg + 7-arg overloads
 * (src/projection/incflo_apply_nodal_projection.cpp:29-267), single level.
 *   velocity (ng ghosts, 3 comps) in/out, velocity_o same shape (used when
 *   incremental or proj_for_small_dt), density (ngd ghosts) or NULL with
 *   constant density ro_0, gp (3 comps, 0 ghosts) in/out, p_nd nodal in/out.
 *   Ghost cells of velocity are zeroed (setBndry), then the caller-supplied
 *   inflow values (inflow_vel, same layout as velocity, may be NULL) are
 *   copied into the first ghost layer of faces with bc == ORC_BC_INFLOW when
 *   set_inflow_bc holds.
 */
int orc_apply_nodal_projection(const orc_params* p, double* velocity, const double* velocity_o, int ng,
                               const double* density, int ngd, double ro_0, double* gp, double* p_nd,
                               const double* inflow_vel, double scaling_factor, int incremental,
                               int proj_for_small_dt, double rtol, double atol, orc_stats* st);

/* ---- building blocks exposed for per-kernel parity tests (unique-node layout) ----
 * Unique-node layout: nn[d] = n[d] (periodic) or n[d]+1, i fastest, no ghosts. */
typedef struct orc_mg orc_mg;
orc_mg* orc_mg_create(const orc_params* p, const double* sigma, double const_sigma);
void    orc_mg_destroy(orc_mg* mg);
int     orc_mg_nlevels(const orc_mg* mg);
void    orc_mg_level_dims(const orc_mg* mg, int lev, int n[3], int nn[3]);
const double* orc_mg_sigma(const orc_mg* mg, int lev);       /* NULL if constant sigma   */
void    orc_adotx(const orc_mg* mg, int lev, const double* phi, double* y);
void    orc_residual(const orc_mg* mg, int lev, const double* phi, const double* rhs, double* res);
void    orc_smooth(const orc_mg* mg, int lev, double* phi, const double* rhs, int nsweeps);
void    orc_restrict(const orc_mg* mg, int flev, const double* fine, double* crse);
void    orc_interp_add(const orc_mg* mg, int flev, double* fine, const double* crse);
void    orc_divu(const orc_mg* mg, const double* vel, int ng, double* rhs);   /* level 0 */
void    orc_mknewu(const orc_mg* mg, const double* phi, double* vel, int ng, double* gphi);
int     orc_bottom_solve(const orc_mg* mg, double* x, const double* b);   /* returns iterations */
int     orc_mlmg_solve(const orc_mg* mg, double* phi, double* rhs, double rtol, double atol, orc_stats* st);
double  orc_dot_weight(const orc_mg* mg, int lev, int i, int j, int k);
void    orc_dot_weights(const orc_mg* mg, int lev, double* w);   /* the whole level, unique-node layout */

#ifdef __cplusplus
}
#endif
#endif
