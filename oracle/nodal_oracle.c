/*
 * nodal_oracle.c -- CPU oracle (test infrastructure, see nodal_oracle.h header:
 * PARITY UNPINNED, never on the product path).
 *
 * Restates, for a single AMR level:
 *   incflo::ApplyNodalProjection   /root/reference/src/projection/incflo_apply_nodal_projection.cpp:29-267
 *   incflo::get_projection_bc      /root/reference/src/projection/incflo_projection_bc.cpp:5-41 (BC codes)
 *   Hydro::NodalProjector::project [U] AMReX-Hydro Projections/hydro_NodalProjector.cpp   
 *   MLNodeLaplacian / MLNodeLinOp  [U] amrex Src/LinearSolvers/MLMG                        
 *   MLMG::solve / mgVcycle         [U] amrex AMReX_MLMG.cpp                                
 *   MLCGSolver BiCGStab / CG       [U] amrex AMReX_MLCGSolver.cpp                          
 *
 * Internal representation: every nodal field is stored once per *unique* node
 * (nn = n in a periodic direction, n+1 otherwise); neighbours across a periodic
 * face wrap, neighbours across a Neumann/inflow face reflect (phi(-1)=phi(1),
 * SURVEY A.8 applyBC), Dirichlet-face nodes are masked (phi = 0).  sigma ghost
 * cells: periodic wrap, otherwise copy of the adjacent interior cell (A.8).
 */
#include "nodal_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXLEV 32

typedef struct {
    int     n[3], nn[3];
    double  dxinv[3];
    double* sigma; /* NULL => constant */
    double  csig;
    long    nnodes, ncells;
    double *cor, *res, *rescor, *old; /* V-cycle work arrays */
} level;

struct orc_mg {
    orc_params p;
    int        nlev;
    int        singular;
    level      L[MAXLEV];
};

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

void orc_default_params(orc_params* p)
{
    memset(p, 0, sizeof(*p));
    for (int d = 0; d < 3; ++d) { p->n[d] = 32; p->dx[d] = 1.0 / 32; p->box[d] = 32; }
    p->max_coarsening_level = 100; /* src/incflo.H:458 */
    p->maxiter = 100; p->bottom_maxiter = 100; p->bottom_rtol = 1e-4; p->bottom_atol = -1.0;
    p->nu1 = 2; p->nu2 = 2; p->nsweeps = 4;
    p->smoother = ORC_SM_LEX; p->box_order = ORC_SM_LEX; p->box_stale_per_call = 1;
}

int orc_set_num_threads(int n)
{
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
}

/* ---------------------------------------------------------------- index maps */
static inline int is_per(const orc_mg* mg, int d) { return mg->p.bclo[d] == ORC_BC_PERIODIC; }
static inline int refl_lo(const orc_mg* mg, int d) { return mg->p.bclo[d] == ORC_BC_NEUMANN || mg->p.bclo[d] == ORC_BC_INFLOW; }
static inline int refl_hi(const orc_mg* mg, int d) { return mg->p.bchi[d] == ORC_BC_NEUMANN || mg->p.bchi[d] == ORC_BC_INFLOW; }

/* node index i in [-1, n+1] -> storage index */
static inline int nmap(const orc_mg* mg, const level* L, int d, int i)
{
    int n = L->n[d];
    if (is_per(mg, d)) { if (i < 0) i += n; else if (i >= n) i -= n; return i; }
    if (i < 0) return -i;        /* reflection; Dirichlet never reads it      */
    if (i > n) return 2 * n - i;
    return i;
}
/* cell index i in [-1, n] -> storage index */
static inline int cmap(const orc_mg* mg, const level* L, int d, int i)
{
    int n = L->n[d];
    if (is_per(mg, d)) { if (i < 0) i += n; else if (i >= n) i -= n; return i; }
    if (i < 0) return 0;
    if (i >= n) return n - 1;
    return i;
}
static inline int masked(const orc_mg* mg, const level* L, int i, int j, int k)
{
    const int idx[3] = {i, j, k};
    for (int d = 0; d < 3; ++d) {
        if (is_per(mg, d)) continue;
        if (idx[d] == 0 && mg->p.bclo[d] == ORC_BC_DIRICHLET) return 1;
        if (idx[d] == L->n[d] && mg->p.bchi[d] == ORC_BC_DIRICHLET) return 1;
        /* mixed face: the outflow part of the overset mask (make_nodalBC_mask), injected to this level */
        if (mg->p.mixed_lo[d] || mg->p.mixed_hi[d]) {
            const int half = mg->p.mix_half / (mg->p.n[mg->p.mix_dir] / L->n[mg->p.mix_dir]);
            if (idx[d] == 0 && mg->p.mixed_lo[d] && idx[mg->p.mix_dir] <= half) return 1;
            if (idx[d] == L->n[d] && mg->p.mixed_hi[d] && idx[mg->p.mix_dir] > half) return 1;
        }
    }
    return 0;
}
#define NIDX(L, i, j, k) (((long)(k) * (L)->nn[1] + (j)) * (L)->nn[0] + (i))
#define CIDX(L, i, j, k) (((long)(k) * (L)->n[1] + (j)) * (L)->n[0] + (i))

/* gather the 8 sigma values around node (i,j,k): S[a][b][c] = sigma(i-1+a, j-1+b, k-1+c) */
static inline void gather_sigma(const orc_mg* mg, const level* L, int i, int j, int k, double S[2][2][2])
{
    if (!L->sigma) {
        for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int c = 0; c < 2; ++c) S[a][b][c] = L->csig;
        return;
    }
    int ci[2] = {cmap(mg, L, 0, i - 1), cmap(mg, L, 0, i)};
    int cj[2] = {cmap(mg, L, 1, j - 1), cmap(mg, L, 1, j)};
    int ck[2] = {cmap(mg, L, 2, k - 1), cmap(mg, L, 2, k)};
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int c = 0; c < 2; ++c)
        S[a][b][c] = L->sigma[CIDX(L, ci[a], cj[b], ck[c])];
}

/* gather 27 phi values; if old != NULL, values outside box (bi,bj,bk) come from old */
static inline void gather_phi(const orc_mg* mg, const level* L, const double* phi, const double* old,
                              const int* bsz, int i, int j, int k, double P[3][3][3])
{
    int ix[3], jy[3], kz[3];
    for (int a = 0; a < 3; ++a) {
        ix[a] = nmap(mg, L, 0, i - 1 + a); jy[a] = nmap(mg, L, 1, j - 1 + a); kz[a] = nmap(mg, L, 2, k - 1 + a);
    }
    if (!old) {
        for (int c = 0; c < 3; ++c) for (int b = 0; b < 3; ++b) for (int a = 0; a < 3; ++a)
            P[a][b][c] = phi[NIDX(L, ix[a], jy[b], kz[c])];
    } else {
        /* bsz[3..5]: number of boxes per direction (the last box may own one more node plane) */
#define BOXOF(idx, d) ((idx) / bsz[d] < bsz[3 + (d)] ? (idx) / bsz[d] : bsz[3 + (d)] - 1)
        int bi = BOXOF(i, 0), bj = BOXOF(j, 1), bk = BOXOF(k, 2);
        int inx[3], iny[3], inz[3];
        /* a neighbour reached through a periodic wrap or a reflection counts as outside the box
         * (its value is the previous sweep's), exactly like a halo cell of the GPU tile */
        for (int a = 0; a < 3; ++a) {
            int ri = i - 1 + a, rj = j - 1 + a, rk = k - 1 + a;
            inx[a] = (ri >= 0 && ri < L->nn[0] && BOXOF(ri, 0) == bi);
            iny[a] = (rj >= 0 && rj < L->nn[1] && BOXOF(rj, 1) == bj);
            inz[a] = (rk >= 0 && rk < L->nn[2] && BOXOF(rk, 2) == bk);
        }
#undef BOXOF
        for (int c = 0; c < 3; ++c) for (int b = 0; b < 3; ++b) for (int a = 0; a < 3; ++a) {
            long id = NIDX(L, ix[a], jy[b], kz[c]);
            P[a][b][c] = (inx[a] && iny[b] && inz[c]) ? phi[id] : old[id];
        }
    }
}

/* y = L phi at one node and the diagonal s0 (mlndlap_adotx_aa / _c) */
static inline double stencil_apply(const level* L, double S[2][2][2], double P[3][3][3], double* s0out)
{
    const double fx = L->dxinv[0] * L->dxinv[0] / 36.0, fy = L->dxinv[1] * L->dxinv[1] / 36.0,
                 fz = L->dxinv[2] * L->dxinv[2] / 36.0;
    const double fxyz = fx + fy + fz, fmx2y2z = -fx + 2 * fy + 2 * fz, f2xmy2z = 2 * fx - fy + 2 * fz,
                 f2x2ymz = 2 * fx + 2 * fy - fz, f4xm2ym2z = 4 * fx - 2 * fy - 2 * fz,
                 fm2x4ym2z = -2 * fx + 4 * fy - 2 * fz, fm2xm2y4z = -2 * fx - 2 * fy + 4 * fz;
    double sumS = 0, corner = 0, ex = 0, ey = 0, ez = 0, fxs = 0, fys = 0, fzs = 0;
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int c = 0; c < 2; ++c) {
        sumS += S[a][b][c];
        corner += S[a][b][c] * P[2 * a][2 * b][2 * c];
    }
    for (int b = 0; b < 2; ++b) for (int c = 0; c < 2; ++c) ex += (S[0][b][c] + S[1][b][c]) * P[1][2 * b][2 * c];
    for (int a = 0; a < 2; ++a) for (int c = 0; c < 2; ++c) ey += (S[a][0][c] + S[a][1][c]) * P[2 * a][1][2 * c];
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) ez += (S[a][b][0] + S[a][b][1]) * P[2 * a][2 * b][1];
    for (int a = 0; a < 2; ++a) fxs += (S[a][0][0] + S[a][0][1] + S[a][1][0] + S[a][1][1]) * P[2 * a][1][1];
    for (int b = 0; b < 2; ++b) fys += (S[0][b][0] + S[0][b][1] + S[1][b][0] + S[1][b][1]) * P[1][2 * b][1];
    for (int c = 0; c < 2; ++c) fzs += (S[0][0][c] + S[0][1][c] + S[1][0][c] + S[1][1][c]) * P[1][1][2 * c];
    double s0 = -4.0 * fxyz * sumS;
    *s0out = s0;
    return s0 * P[1][1][1] + fxyz * corner + fmx2y2z * ex + f2xmy2z * ey + f2x2ymz * ez + f4xm2ym2z * fxs +
           fm2x4ym2z * fys + fm2xm2y4z * fzs;
}

static inline double node_Lphi(const orc_mg* mg, const level* L, const double* phi, const double* old,
                               const int* bsz, int i, int j, int k, double* s0)
{
    double S[2][2][2], P[3][3][3];
    gather_sigma(mg, L, i, j, k, S);
    gather_phi(mg, L, phi, old, bsz, i, j, k, P);
    return stencil_apply(L, S, P, s0);
}

/* ---------------------------------------------------------------- hierarchy */
orc_mg* orc_mg_create(const orc_params* p, const double* sigma, double const_sigma)
{
    orc_mg* mg = (orc_mg*)calloc(1, sizeof(orc_mg));
    mg->p = *p;
    mg->singular = 1;
    for (int d = 0; d < 3; ++d) {
        if (p->bclo[d] == ORC_BC_DIRICHLET || p->bchi[d] == ORC_BC_DIRICHLET) mg->singular = 0;
        if (p->mixed_lo[d] || p->mixed_hi[d]) mg->singular = 0;
    }
    int n[3] = {p->n[0], p->n[1], p->n[2]};
    double dx[3] = {p->dx[0], p->dx[1], p->dx[2]};
    int lev = 0;
    for (;;) {
        level* L = &mg->L[lev];
        L->nnodes = 1; L->ncells = 1;
        for (int d = 0; d < 3; ++d) {
            L->n[d] = n[d]; L->nn[d] = n[d] + (is_per(mg, d) ? 0 : 1); L->dxinv[d] = 1.0 / dx[d];
            L->nnodes *= L->nn[d]; L->ncells *= n[d];
        }
        L->csig = const_sigma;
        if (sigma) {
            L->sigma = (double*)malloc(sizeof(double) * L->ncells);
            if (lev == 0) memcpy(L->sigma, sigma, sizeof(double) * L->ncells);
            else { /* average_down, arithmetic mean of the 8 children (A.8) */
                const level* F = &mg->L[lev - 1];
                for (int k = 0; k < n[2]; ++k) for (int j = 0; j < n[1]; ++j) for (int i = 0; i < n[0]; ++i) {
                    double s = 0;
                    for (int c = 0; c < 2; ++c) for (int b = 0; b < 2; ++b) for (int a = 0; a < 2; ++a)
                        s += F->sigma[CIDX(F, 2 * i + a, 2 * j + b, 2 * k + c)];
                    L->sigma[CIDX(L, i, j, k)] = 0.125 * s;
                }
            }
        }
        L->cor = (double*)calloc(L->nnodes, sizeof(double));
        L->res = (double*)calloc(L->nnodes, sizeof(double));
        L->rescor = (double*)calloc(L->nnodes, sizeof(double));
        L->old = (double*)calloc(L->nnodes, sizeof(double));
        ++lev;
        /* coarsen by 2 while every direction stays even and >= 2 cells wide (A.8, mg_box_min_width=2) */
        int ok = (lev <= p->max_coarsening_level) && lev < MAXLEV;
        for (int d = 0; d < 3; ++d) if (n[d] % 2 != 0 || n[d] / 2 < 2) ok = 0;
        if (!ok) break;
        for (int d = 0; d < 3; ++d) { n[d] /= 2; dx[d] *= 2; }
    }
    mg->nlev = lev;
    return mg;
}

void orc_mg_destroy(orc_mg* mg)
{
    if (!mg) return;
    for (int l = 0; l < mg->nlev; ++l) {
        free(mg->L[l].sigma); free(mg->L[l].cor); free(mg->L[l].res); free(mg->L[l].rescor); free(mg->L[l].old);
    }
    free(mg);
}
int orc_mg_nlevels(const orc_mg* mg) { return mg->nlev; }
void orc_mg_level_dims(const orc_mg* mg, int lev, int n[3], int nn[3])
{
    for (int d = 0; d < 3; ++d) { n[d] = mg->L[lev].n[d]; nn[d] = mg->L[lev].nn[d]; }
}
const double* orc_mg_sigma(const orc_mg* mg, int lev) { return mg->L[lev].sigma; }

/* dot-product weight: 1/2 per Neumann/inflow boundary direction the node lies on (A.8) */
double orc_dot_weight(const orc_mg* mg, int lev, int i, int j, int k)
{
    const level* L = &mg->L[lev];
    const int idx[3] = {i, j, k};
    double w = 1.0;
    if (masked(mg, L, i, j, k)) return 0.0;
    for (int d = 0; d < 3; ++d) {
        if (is_per(mg, d)) continue;
        if (idx[d] == 0 && refl_lo(mg, d)) w *= 0.5;
        if (idx[d] == L->n[d] && refl_hi(mg, d)) w *= 0.5;
    }
    return w;
}

/* ---------------------------------------------------------------- operator */
void orc_adotx(const orc_mg* mg, int lev, const double* phi, double* y)
{
    const level* L = &mg->L[lev];
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < L->nn[2]; ++k) for (int j = 0; j < L->nn[1]; ++j) for (int i = 0; i < L->nn[0]; ++i) {
        double s0;
        y[NIDX(L, i, j, k)] = masked(mg, L, i, j, k) ? 0.0 : node_Lphi(mg, L, phi, NULL, NULL, i, j, k, &s0);
    }
}

void orc_residual(const orc_mg* mg, int lev, const double* phi, const double* rhs, double* res)
{
    const level* L = &mg->L[lev];
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < L->nn[2]; ++k) for (int j = 0; j < L->nn[1]; ++j) for (int i = 0; i < L->nn[0]; ++i) {
        double s0;
        long id = NIDX(L, i, j, k);
        res[id] = masked(mg, L, i, j, k) ? 0.0 : rhs[id] - node_Lphi(mg, L, phi, NULL, NULL, i, j, k, &s0);
    }
}

/* ---------------------------------------------------------------- smoothers (A.4) */
static inline void gs_node(const orc_mg* mg, const level* L, double* phi, const double* old, const int* bsz,
                           const double* rhs, int i, int j, int k)
{
    long id = NIDX(L, i, j, k);
    if (masked(mg, L, i, j, k)) { phi[id] = 0.0; return; }
    double s0;
    double Ax = node_Lphi(mg, L, phi, old, bsz, i, j, k, &s0);
    phi[id] += (rhs[id] - Ax) / s0;
}

static void sweep_ordered(const orc_mg* mg, const level* L, double* phi, const double* old, const int* bsz,
                          const double* rhs, int order, int lo0, int hi0, int lo1, int hi1, int lo2, int hi2,
                          int par)
{
    if (order == ORC_SM_LEX) {
        for (int k = lo2; k < hi2; ++k) for (int j = lo1; j < hi1; ++j) for (int i = lo0; i < hi0; ++i)
            gs_node(mg, L, phi, old, bsz, rhs, i, j, k);
    } else if (order == ORC_SM_COLOR8) {
        for (int color = 0; color < 8; ++color) {
            int pi = color & 1, pj = (color >> 1) & 1, pk = (color >> 2) & 1;
#pragma omp parallel for collapse(2) schedule(static) if (par)
            for (int k = lo2; k < hi2; ++k) for (int j = lo1; j < hi1; ++j) {
                if ((k & 1) != pk || (j & 1) != pj) continue;
                for (int i = lo0 + (((lo0 & 1) != pi) ? 1 : 0); i < hi0; i += 2) gs_node(mg, L, phi, old, bsz, rhs, i, j, k);
            }
        }
    } else if (order == ORC_SM_PLANE4) { /* plane by plane (k ascending); inside a plane 4 colours c=(i&1)+2(j&1) */
        for (int k = lo2; k < hi2; ++k) for (int color = 0; color < 4; ++color) {
            int pi = color & 1, pj = (color >> 1) & 1;
#pragma omp parallel for schedule(static) if (par)
            for (int j = lo1; j < hi1; ++j) {
                if ((j & 1) != pj) continue;
                for (int i = lo0 + (((lo0 & 1) != pi) ? 1 : 0); i < hi0; i += 2) gs_node(mg, L, phi, old, bsz, rhs, i, j, k);
            }
        }
    } else { /* COLOR4XY: planes swept bottom to top inside a colour (matters only when dz != dx,dy) */
        for (int color = 0; color < 4; ++color) {
            int pi = color & 1, pj = (color >> 1) & 1;
            for (int k = lo2; k < hi2; ++k) {
#pragma omp parallel for schedule(static) if (par)
                for (int j = lo1; j < hi1; ++j) {
                    if ((j & 1) != pj) continue;
                    for (int i = lo0 + (((lo0 & 1) != pi) ? 1 : 0); i < hi0; i += 2) gs_node(mg, L, phi, old, bsz, rhs, i, j, k);
                }
            }
        }
    }
}

static void smooth_level(const orc_mg* mg, level* L, double* phi, const double* rhs, int nsweeps)
{
    const orc_params* p = &mg->p;
    if (p->smoother == ORC_SM_JACOBI) {
        for (int s = 0; s < nsweeps; ++s) {
            memcpy(L->old, phi, sizeof(double) * L->nnodes);
#pragma omp parallel for collapse(2) schedule(static)
            for (int k = 0; k < L->nn[2]; ++k) for (int j = 0; j < L->nn[1]; ++j) for (int i = 0; i < L->nn[0]; ++i) {
                long id = NIDX(L, i, j, k);
                if (masked(mg, L, i, j, k)) { phi[id] = 0.0; continue; }
                double s0, Ax = node_Lphi(mg, L, L->old, NULL, NULL, i, j, k, &s0);
                phi[id] = L->old[id] + (2.0 / 3.0) * (rhs[id] - Ax) / s0;
            }
        }
        return;
    }
    if (p->smoother == ORC_SM_BOX) {
        /* z-chunk rule shared with the CUDA smoother: aim at 592 tile-chunks per sweep (148 SMs x 2
         * resident CTAs x 2 waves), chunk height between 4 and the cap box[2] */
        int bsz[6] = {p->box[0], p->box[1], p->box[2], 0, 0, 0};
        int* nb = bsz + 3;
        if (p->box_amrex) {   /* the reference's grids: max_grid_size boxes, coarsened with the level */
            int lev = (int)(L - mg->L);
            for (int d = 0; d < 3; ++d) {
                int b = p->box[d] >> lev; if (b < 2) b = 2;
                if (b > L->n[d]) b = L->n[d];
                bsz[d] = b;
                nb[d] = (L->n[d] + b - 1) / b;   /* cell boxes; the last one also owns the top node plane */
            }
        } else {
            int ntiles = ((L->nn[0] + bsz[0] - 1) / bsz[0]) * ((L->nn[1] + bsz[1] - 1) / bsz[1]);
            int nch = 592 / ntiles; if (nch < 1) nch = 1;
            int c = (L->nn[2] + nch - 1) / nch;
            if (c < bsz[2]) bsz[2] = c;
            /* lower bound of the chunk height: 8 planes on levels with more than 33 node planes (shorter
             * chunks cost a V-cycle there: measured 4 -> 8 cycles, 8 -> 7); coarse levels are insensitive
             * (same residual history to 2 digits), so they use 4 / 2 / 1 to expose more parallel chunks */
            { int mn = L->nn[2] > 33 ? 8 : L->nn[2] > 17 ? 4 : L->nn[2] > 9 ? 2 : 1;
              if (bsz[2] < mn) bsz[2] = mn; }
            for (int d = 0; d < 3; ++d) nb[d] = (L->nn[d] + bsz[d] - 1) / bsz[d];
        }
        int outer = p->box_stale_per_call ? 1 : nsweeps, inner = p->box_stale_per_call ? nsweeps : 1;
        for (int so = 0; so < outer; ++so) {
            memcpy(L->old, phi, sizeof(double) * L->nnodes);
#pragma omp parallel for collapse(3) schedule(dynamic)
            for (int bk = 0; bk < nb[2]; ++bk) for (int bj = 0; bj < nb[1]; ++bj) for (int bi = 0; bi < nb[0]; ++bi) {
                int lo0 = bi * bsz[0], lo1 = bj * bsz[1], lo2 = bk * bsz[2];
                int hi0 = (bi + 1 < nb[0] && lo0 + bsz[0] < L->nn[0]) ? lo0 + bsz[0] : L->nn[0];
                int hi1 = (bj + 1 < nb[1] && lo1 + bsz[1] < L->nn[1]) ? lo1 + bsz[1] : L->nn[1];
                int hi2 = (bk + 1 < nb[2] && lo2 + bsz[2] < L->nn[2]) ? lo2 + bsz[2] : L->nn[2];
                for (int si = 0; si < inner; ++si)
                    sweep_ordered(mg, L, phi, L->old, bsz, rhs, p->box_order, lo0, hi0, lo1, hi1, lo2, hi2, 0);
            }
        }
        return;
    }
    for (int s = 0; s < nsweeps; ++s)
        sweep_ordered(mg, L, phi, NULL, NULL, rhs, p->smoother, 0, L->nn[0], 0, L->nn[1], 0, L->nn[2], 1);
}

void orc_smooth(const orc_mg* mg, int lev, double* phi, const double* rhs, int nsweeps)
{
    smooth_level(mg, (level*)&mg->L[lev], phi, rhs, nsweeps);
}

/* ---------------------------------------------------------------- restriction (A.5) */
void orc_restrict(const orc_mg* mg, int flev, const double* fine, double* crse)
{
    const level *F = &mg->L[flev], *C = &mg->L[flev + 1];
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < C->nn[2]; ++k) for (int j = 0; j < C->nn[1]; ++j) for (int i = 0; i < C->nn[0]; ++i) {
        long id = NIDX(C, i, j, k);
        if (masked(mg, C, i, j, k)) { crse[id] = 0.0; continue; }
        double s = 0;
        for (int c = -1; c <= 1; ++c) for (int b = -1; b <= 1; ++b) for (int a = -1; a <= 1; ++a) {
            double w = (a ? 1.0 : 2.0) * (b ? 1.0 : 2.0) * (c ? 1.0 : 2.0);
            s += w * fine[NIDX(F, nmap(mg, F, 0, 2 * i + a), nmap(mg, F, 1, 2 * j + b), nmap(mg, F, 2, 2 * k + c))];
        }
        crse[id] = s / 64.0;
    }
}

/* ---------------------------------------------------------------- interpolation (A.6) */
typedef struct { const orc_mg* mg; const level *F, *C; const double* crse; } ictx;

static inline double sg(const ictx* x, int i, int j, int k)
{
    const level* F = x->F;
    if (!F->sigma) return F->csig;
    return F->sigma[CIDX(F, cmap(x->mg, F, 0, i), cmap(x->mg, F, 1, j), cmap(x->mg, F, 2, k))];
}
static inline double cr(const ictx* x, int ic, int jc, int kc)
{
    const level* C = x->C;
    return x->crse[NIDX(C, nmap(x->mg, C, 0, ic), nmap(x->mg, C, 1, jc), nmap(x->mg, C, 2, kc))];
}
/* quad sums of sigma on the low/high side of fine node (i,j,k) in each direction */
static inline double qx(const ictx* x, int i, int j, int k, int side) /* side 0: cells i-1, 1: cells i */
{
    int ii = i - 1 + side;
    return sg(x, ii, j - 1, k - 1) + sg(x, ii, j, k - 1) + sg(x, ii, j - 1, k) + sg(x, ii, j, k);
}
static inline double qy(const ictx* x, int i, int j, int k, int side)
{
    int jj = j - 1 + side;
    return sg(x, i - 1, jj, k - 1) + sg(x, i, jj, k - 1) + sg(x, i - 1, jj, k) + sg(x, i, jj, k);
}
static inline double qz(const ictx* x, int i, int j, int k, int side)
{
    int kk = k - 1 + side;
    return sg(x, i - 1, j - 1, kk) + sg(x, i, j - 1, kk) + sg(x, i - 1, j, kk) + sg(x, i, j, kk);
}
/* (i,j,k) fine node; (ic,jc,kc) = coarse node at/below it in the odd directions */
static double line_x(const ictx* x, int i, int j, int k, int ic, int jc, int kc)
{
    double w1 = qx(x, i, j, k, 0), w2 = qx(x, i, j, k, 1);
    return (w1 * cr(x, ic, jc, kc) + w2 * cr(x, ic + 1, jc, kc)) / (w1 + w2);
}
static double line_y(const ictx* x, int i, int j, int k, int ic, int jc, int kc)
{
    double w1 = qy(x, i, j, k, 0), w2 = qy(x, i, j, k, 1);
    return (w1 * cr(x, ic, jc, kc) + w2 * cr(x, ic, jc + 1, kc)) / (w1 + w2);
}
static double line_z(const ictx* x, int i, int j, int k, int ic, int jc, int kc)
{
    double w1 = qz(x, i, j, k, 0), w2 = qz(x, i, j, k, 1);
    return (w1 * cr(x, ic, jc, kc) + w2 * cr(x, ic, jc, kc + 1)) / (w1 + w2);
}
static double face_xy(const ictx* x, int i, int j, int k, int ic, int jc, int kc)
{
    double w1 = qx(x, i, j, k, 0), w2 = qx(x, i, j, k, 1), w3 = qy(x, i, j, k, 0), w4 = qy(x, i, j, k, 1);
    return (w1 * line_y(x, i - 1, j, k, ic, jc, kc) + w2 * line_y(x, i + 1, j, k, ic + 1, jc, kc) +
            w3 * line_x(x, i, j - 1, k, ic, jc, kc) + w4 * line_x(x, i, j + 1, k, ic, jc + 1, kc)) / (w1 + w2 + w3 + w4);
}
static double face_xz(const ictx* x, int i, int j, int k, int ic, int jc, int kc)
{
    double w1 = qx(x, i, j, k, 0), w2 = qx(x, i, j, k, 1), w3 = qz(x, i, j, k, 0), w4 = qz(x, i, j, k, 1);
    return (w1 * line_z(x, i - 1, j, k, ic, jc, kc) + w2 * line_z(x, i + 1, j, k, ic + 1, jc, kc) +
            w3 * line_x(x, i, j, k - 1, ic, jc, kc) + w4 * line_x(x, i, j, k + 1, ic, jc, kc + 1)) / (w1 + w2 + w3 + w4);
}
static double face_yz(const ictx* x, int i, int j, int k, int ic, int jc, int kc)
{
    double w1 = qy(x, i, j, k, 0), w2 = qy(x, i, j, k, 1), w3 = qz(x, i, j, k, 0), w4 = qz(x, i, j, k, 1);
    return (w1 * line_z(x, i, j - 1, k, ic, jc, kc) + w2 * line_z(x, i, j + 1, k, ic, jc + 1, kc) +
            w3 * line_y(x, i, j, k - 1, ic, jc, kc) + w4 * line_y(x, i, j, k + 1, ic, jc, kc + 1)) / (w1 + w2 + w3 + w4);
}

void orc_interp_add(const orc_mg* mg, int flev, double* fine, const double* crse)
{
    const level *F = &mg->L[flev], *C = &mg->L[flev + 1];
    ictx x = {mg, F, C, crse};
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < F->nn[2]; ++k) for (int j = 0; j < F->nn[1]; ++j) for (int i = 0; i < F->nn[0]; ++i) {
        if (masked(mg, F, i, j, k)) continue;
        int ic = i >> 1, jc = j >> 1, kc = k >> 1;
        int io = i & 1, jo = j & 1, ko = k & 1;
        double v;
        if (io && jo && ko) {
            double w1 = qx(&x, i, j, k, 0), w2 = qx(&x, i, j, k, 1), w3 = qy(&x, i, j, k, 0), w4 = qy(&x, i, j, k, 1),
                   w5 = qz(&x, i, j, k, 0), w6 = qz(&x, i, j, k, 1);
            v = (w1 * face_yz(&x, i - 1, j, k, ic, jc, kc) + w2 * face_yz(&x, i + 1, j, k, ic + 1, jc, kc) +
                 w3 * face_xz(&x, i, j - 1, k, ic, jc, kc) + w4 * face_xz(&x, i, j + 1, k, ic, jc + 1, kc) +
                 w5 * face_xy(&x, i, j, k - 1, ic, jc, kc) + w6 * face_xy(&x, i, j, k + 1, ic, jc, kc + 1)) /
                (w1 + w2 + w3 + w4 + w5 + w6);
        } else if (jo && ko) v = face_yz(&x, i, j, k, ic, jc, kc);
        else if (io && ko)   v = face_xz(&x, i, j, k, ic, jc, kc);
        else if (io && jo)   v = face_xy(&x, i, j, k, ic, jc, kc);
        else if (io)         v = line_x(&x, i, j, k, ic, jc, kc);
        else if (jo)         v = line_y(&x, i, j, k, ic, jc, kc);
        else if (ko)         v = line_z(&x, i, j, k, ic, jc, kc);
        else                 v = cr(&x, ic, jc, kc);
        fine[NIDX(F, i, j, k)] += v;
    }
}

/* ---------------------------------------------------------------- rhs (A.2) */
static inline double velg(const orc_mg* mg, const level* L, const double* vel, int ng, int comp, int i, int j, int k)
{
    /* periodic ghosts: wrap (vel.FillBoundary); others: read the caller's ghost cell */
    if (is_per(mg, 0)) i = cmap(mg, L, 0, i);
    if (is_per(mg, 1)) j = cmap(mg, L, 1, j);
    if (is_per(mg, 2)) k = cmap(mg, L, 2, k);
    long sx = L->n[0] + 2 * ng, sy = L->n[1] + 2 * ng, sz = L->n[2] + 2 * ng;
    return vel[(((long)comp * sz + (k + ng)) * sy + (j + ng)) * sx + (i + ng)];
}

void orc_divu(const orc_mg* mg, const double* vel, int ng, double* rhs)
{
    const level* L = &mg->L[0];
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < L->nn[2]; ++k) for (int j = 0; j < L->nn[1]; ++j) for (int i = 0; i < L->nn[0]; ++i) {
        long id = NIDX(L, i, j, k);
        if (masked(mg, L, i, j, k)) { rhs[id] = 0.0; continue; }
        const int idx[3] = {i, j, k};
        double z[3][2]; /* z[d][0]: weight of the low-side cells (idx-1), z[d][1]: high side */
        double scale = 1.0;
        for (int d = 0; d < 3; ++d) {
            z[d][0] = z[d][1] = 1.0;
            if (is_per(mg, d)) continue;
            if (idx[d] == 0 && refl_lo(mg, d)) { z[d][0] = 0.0; scale *= 2.0; }
            if (idx[d] == L->n[d] && refl_hi(mg, d)) { z[d][1] = 0.0; scale *= 2.0; }
        }
        double dux = 0, dvy = 0, dwz = 0;
        for (int c = 0; c < 2; ++c) for (int b = 0; b < 2; ++b) {
            dux += (velg(mg, L, vel, ng, 0, i, j - 1 + b, k - 1 + c) - velg(mg, L, vel, ng, 0, i - 1, j - 1 + b, k - 1 + c)) * z[1][b] * z[2][c];
            dvy += (velg(mg, L, vel, ng, 1, i - 1 + b, j, k - 1 + c) - velg(mg, L, vel, ng, 1, i - 1 + b, j - 1, k - 1 + c)) * z[0][b] * z[2][c];
            dwz += (velg(mg, L, vel, ng, 2, i - 1 + b, j - 1 + c, k) - velg(mg, L, vel, ng, 2, i - 1 + b, j - 1 + c, k - 1)) * z[0][b] * z[1][c];
        }
        rhs[id] = scale * 0.25 * (L->dxinv[0] * dux + L->dxinv[1] * dvy + L->dxinv[2] * dwz);
    }
}

/* ---------------------------------------------------------------- velocity update + gradient (A.7) */
void orc_mknewu(const orc_mg* mg, const double* phi, double* vel, int ng, double* gphi)
{
    const level* L = &mg->L[0];
    long sx = L->n[0] + 2 * ng, sy = L->n[1] + 2 * ng, sz = L->n[2] + 2 * ng;
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < L->n[2]; ++k) for (int j = 0; j < L->n[1]; ++j) for (int i = 0; i < L->n[0]; ++i) {
        double P[2][2][2];
        for (int c = 0; c < 2; ++c) for (int b = 0; b < 2; ++b) for (int a = 0; a < 2; ++a)
            P[a][b][c] = phi[NIDX(L, nmap(mg, L, 0, i + a), nmap(mg, L, 1, j + b), nmap(mg, L, 2, k + c))];
        double gx = 0.25 * L->dxinv[0] * ((P[1][0][0] - P[0][0][0]) + (P[1][1][0] - P[0][1][0]) + (P[1][0][1] - P[0][0][1]) + (P[1][1][1] - P[0][1][1]));
        double gy = 0.25 * L->dxinv[1] * ((P[0][1][0] - P[0][0][0]) + (P[1][1][0] - P[1][0][0]) + (P[0][1][1] - P[0][0][1]) + (P[1][1][1] - P[1][0][1]));
        double gz = 0.25 * L->dxinv[2] * ((P[0][0][1] - P[0][0][0]) + (P[1][0][1] - P[1][0][0]) + (P[0][1][1] - P[0][1][0]) + (P[1][1][1] - P[1][1][0]));
        double s = L->sigma ? L->sigma[CIDX(L, i, j, k)] : L->csig;
        long c0 = CIDX(L, i, j, k);
        double g[3] = {gx, gy, gz};
        for (int comp = 0; comp < 3; ++comp) {
            long vi = (((long)comp * sz + (k + ng)) * sy + (j + ng)) * sx + (i + ng);
            if (vel) vel[vi] += -s * g[comp];     /* vel += fluxes, fluxes = -sigma G phi */
            if (gphi) gphi[(long)comp * L->ncells + c0] = g[comp];
        }
    }
}

/* ---------------------------------------------------------------- reductions */
static double norminf(const level* L, const double* x)
{
    double m = 0;
#pragma omp parallel for reduction(max : m) schedule(static)
    for (long i = 0; i < L->nnodes; ++i) { double a = fabs(x[i]); if (a > m) m = a; }
    return m;
}
static double wdot(const orc_mg* mg, int lev, const double* x, const double* y)
{
    const level* L = &mg->L[lev];
    double s = 0;
    for (int k = 0; k < L->nn[2]; ++k) for (int j = 0; j < L->nn[1]; ++j) for (int i = 0; i < L->nn[0]; ++i) {
        long id = NIDX(L, i, j, k);
        s += orc_dot_weight(mg, lev, i, j, k) * x[id] * y[id];
    }
    return s;
}
void orc_dot_weights(const orc_mg* mg, int lev, double* w)
{
    const level* L = &mg->L[lev];
    for (int k = 0; k < L->nn[2]; ++k) for (int j = 0; j < L->nn[1]; ++j) for (int i = 0; i < L->nn[0]; ++i)
        w[NIDX(L, i, j, k)] = orc_dot_weight(mg, lev, i, j, k);
}
/* subtract the weighted mean (singular problems, A.8) */
static void make_solvable(const orc_mg* mg, int lev, double* rhs)
{
    if (!mg->singular) return;
    const level* L = &mg->L[lev];
    double s = 0, w = 0;
    for (int k = 0; k < L->nn[2]; ++k) for (int j = 0; j < L->nn[1]; ++j) for (int i = 0; i < L->nn[0]; ++i) {
        double ww = orc_dot_weight(mg, lev, i, j, k);
        s += ww * rhs[NIDX(L, i, j, k)]; w += ww;
    }
    double off = s / w;
    for (long i = 0; i < L->nnodes; ++i) rhs[i] -= off;
}

/* ---------------------------------------------------------------- bottom solvers (A.10) */
static int bicgstab(const orc_mg* mg, int lev, double* x, const double* b, int* iters)
{
    const level* L = &mg->L[lev];
    long n = L->nnodes;
    double *r = (double*)malloc(7 * n * sizeof(double)), *rh = r + n, *p = rh + n, *v = p + n, *s = v + n, *t = s + n;
    memset(x, 0, n * sizeof(double));
    memcpy(r, b, n * sizeof(double)); memcpy(rh, r, n * sizeof(double));
    double rnorm0 = norminf(L, r), rnorm = rnorm0;
    double rtol = mg->p.bottom_rtol, atol = mg->p.bottom_atol;
    int ret = 0, it = 0;
    double rho1 = 0, alpha = 0, omega = 0;
    if (rnorm0 == 0 || rnorm0 < atol) { free(r); *iters = 0; return 0; }
    for (it = 1; it <= mg->p.bottom_maxiter; ++it) {
        double rho = wdot(mg, lev, rh, r);
        if (rho == 0) { ret = 1; break; }
        if (it == 1) memcpy(p, r, n * sizeof(double));
        else {
            double beta = (rho / rho1) * (alpha / omega);
            for (long i = 0; i < n; ++i) p[i] = r[i] + beta * (p[i] - omega * v[i]);
        }
        orc_adotx(mg, lev, p, v);
        double rhTv = wdot(mg, lev, rh, v);
        if (rhTv == 0) { ret = 3; break; }
        alpha = rho / rhTv;
        for (long i = 0; i < n; ++i) { x[i] += alpha * p[i]; s[i] = r[i] - alpha * v[i]; }
        rnorm = norminf(L, s);
        if (rnorm < rtol * rnorm0 || rnorm < atol) break;
        orc_adotx(mg, lev, s, t);
        double tt = wdot(mg, lev, t, t), ts = wdot(mg, lev, t, s);
        if (tt == 0) { ret = 4; break; }
        omega = ts / tt;
        for (long i = 0; i < n; ++i) { x[i] += omega * s[i]; r[i] = s[i] - omega * t[i]; }
        rnorm = norminf(L, r);
        if (rnorm < rtol * rnorm0 || rnorm < atol) break;
        if (omega == 0) { ret = 4; break; }
        rho1 = rho;
    }
    if (ret == 0 && it > mg->p.bottom_maxiter && !(rnorm < rtol * rnorm0)) ret = 8;
    *iters = it > mg->p.bottom_maxiter ? mg->p.bottom_maxiter : it;
    free(r);
    return ret;
}

static int cg(const orc_mg* mg, int lev, double* x, const double* b, int* iters)
{
    const level* L = &mg->L[lev];
    long n = L->nnodes;
    double *r = (double*)malloc(3 * n * sizeof(double)), *p = r + n, *q = p + n;
    memset(x, 0, n * sizeof(double));
    memcpy(r, b, n * sizeof(double));
    double rnorm0 = norminf(L, r), rnorm = rnorm0, rho1 = 0;
    int ret = 0, it = 0;
    if (rnorm0 == 0) { free(r); *iters = 0; return 0; }
    for (it = 1; it <= mg->p.bottom_maxiter; ++it) {
        double rho = wdot(mg, lev, r, r);
        if (rho == 0) { ret = 1; break; }
        if (it == 1) memcpy(p, r, n * sizeof(double));
        else { double beta = rho / rho1; for (long i = 0; i < n; ++i) p[i] = r[i] + beta * p[i]; }
        orc_adotx(mg, lev, p, q);
        double pq = wdot(mg, lev, p, q);
        if (pq == 0) { ret = 1; break; }
        double alpha = rho / pq;
        for (long i = 0; i < n; ++i) { x[i] += alpha * p[i]; r[i] -= alpha * q[i]; }
        rnorm = norminf(L, r);
        if (rnorm < mg->p.bottom_rtol * rnorm0 || rnorm < mg->p.bottom_atol) break;
        rho1 = rho;
    }
    if (ret == 0 && it > mg->p.bottom_maxiter) ret = 8;
    *iters = it > mg->p.bottom_maxiter ? mg->p.bottom_maxiter : it;
    free(r);
    return ret;
}

/* MLMG::actualBottomSolve with bottom_solver = bicgcg: BiCGStab, CG retry, then smoothing fallback */
int orc_bottom_solve(const orc_mg* mg, double* x, const double* b)
{
    int lev = mg->nlev - 1, iters = 0;
    level* L = (level*)&mg->L[lev];
    double* bb = (double*)malloc(L->nnodes * sizeof(double));
    memcpy(bb, b, L->nnodes * sizeof(double));
    make_solvable(mg, lev, bb);
    int ret = bicgstab(mg, lev, x, bb, &iters);
    if (ret != 0) { int it2 = 0; ret = cg(mg, lev, x, bb, &it2); iters += it2; }
    if (ret != 0) {
        memset(x, 0, L->nnodes * sizeof(double));
        for (int i = 0; i < 8; ++i) smooth_level(mg, L, x, bb, mg->p.nsweeps);
    }
    free(bb);
    return iters;
}

/* ---------------------------------------------------------------- MLMG (A.9) */
static void vcycle(const orc_mg* mg, orc_stats* st)
{
    int nl = mg->nlev;
    for (int l = 0; l < nl - 1; ++l) {
        level* L = (level*)&mg->L[l];
        memset(L->cor, 0, L->nnodes * sizeof(double));
        for (int i = 0; i < mg->p.nu1; ++i) smooth_level(mg, L, L->cor, L->res, mg->p.nsweeps);
        orc_residual(mg, l, L->cor, L->res, L->rescor);
        orc_restrict(mg, l, L->rescor, mg->L[l + 1].res);
    }
    {
        level* B = (level*)&mg->L[nl - 1];
        if (nl == 1) { /* single level: the "bottom" solve is all there is */ }
        st->bottom_iters += orc_bottom_solve(mg, B->cor, B->res);
    }
    for (int l = nl - 2; l >= 0; --l) {
        level* L = (level*)&mg->L[l];
        orc_interp_add(mg, l, L->cor, mg->L[l + 1].cor);
        for (int i = 0; i < mg->p.nu2; ++i) smooth_level(mg, L, L->cor, L->res, mg->p.nsweeps);
    }
}

int orc_mlmg_solve(const orc_mg* mg, double* phi, double* rhs, double rtol, double atol, orc_stats* st)
{
    const level* L0 = &mg->L[0];
    double t0 = now_s();
    st->iters = 0; st->bottom_iters = 0; st->status = 0; st->nlevels = mg->nlev;
    /* Dirichlet nodes carry phi = 0, rhs = 0 */
    for (int k = 0; k < L0->nn[2]; ++k) for (int j = 0; j < L0->nn[1]; ++j) for (int i = 0; i < L0->nn[0]; ++i)
        if (masked(mg, L0, i, j, k)) { phi[NIDX(L0, i, j, k)] = 0; rhs[NIDX(L0, i, j, k)] = 0; }
    make_solvable(mg, 0, rhs);
    st->rhsnorm = norminf(L0, rhs);
    orc_residual(mg, 0, phi, rhs, L0->res);
    st->resnorm0 = norminf(L0, L0->res);
    double maxnorm = st->rhsnorm > st->resnorm0 ? st->rhsnorm : st->resnorm0;
    double target = fmax(atol, fmax(rtol, 1e-16) * maxnorm);
    st->resnorm = st->resnorm0;
    st->resnorm_hist[0] = st->resnorm0;
    if (st->resnorm0 <= target) { st->t_solve = now_s() - t0; return 0; }
    int converged = 0;
    for (int it = 0; it < mg->p.maxiter; ++it) {
        vcycle(mg, st);
        for (long i = 0; i < L0->nnodes; ++i) phi[i] += L0->cor[i];
        orc_residual(mg, 0, phi, rhs, L0->res);
        st->resnorm = norminf(L0, L0->res);
        st->iters = it + 1;
        if (it + 1 < 128) st->resnorm_hist[it + 1] = st->resnorm;
        if (mg->p.verbose) printf("oracle MLMG: iter %d resid/bnorm = %.6e\n", it + 1, st->resnorm / maxnorm);
        if (st->resnorm <= target) { converged = 1; break; }
        if (st->resnorm > 1e20 * maxnorm) { st->status = 2; break; }
    }
    if (!converged && st->status == 0) st->status = 1;
    st->t_solve = now_s() - t0;
    return st->status;
}

/* ---------------------------------------------------------------- NodalProjector::project (A.1) */
static void unique_to_amrex_nodal(const orc_mg* mg, const double* u, double* out)
{
    const level* L = &mg->L[0];
    long sx = L->n[0] + 1, sy = L->n[1] + 1;
    for (int k = 0; k <= L->n[2]; ++k) for (int j = 0; j <= L->n[1]; ++j) for (int i = 0; i <= L->n[0]; ++i)
        out[((long)k * sy + j) * sx + i] = u[NIDX(L, nmap(mg, L, 0, i), nmap(mg, L, 1, j), nmap(mg, L, 2, k))];
}

int orc_project(const orc_params* p, double* vel, int ng, const double* sigma, double const_sigma, double* phi,
                double* gphi, double* rhs_out, double rtol, double atol, orc_stats* st)
{
    double t0 = now_s();
    orc_mg* mg = orc_mg_create(p, sigma, const_sigma);
    const level* L = &mg->L[0];
    double* rhs = (double*)calloc(L->nnodes, sizeof(double));
    double* ph = (double*)calloc(L->nnodes, sizeof(double));
    orc_divu(mg, vel, ng, rhs);
    if (rhs_out) unique_to_amrex_nodal(mg, rhs, rhs_out);
    int status = orc_mlmg_solve(mg, ph, rhs, rtol, atol, st);
    orc_mknewu(mg, ph, vel, ng, gphi);
    unique_to_amrex_nodal(mg, ph, phi);
    free(rhs); free(ph);
    orc_mg_destroy(mg);
    st->t_total = now_s() - t0;
    return status;
}

/* ---------------------------------------------------------------- incflo::ApplyNodalProjection */
int orc_apply_nodal_projection(const orc_params* p, double* velocity, const double* velocity_o, int ng,
                               const double* density, int ngd, double ro_0, double* gp, double* p_nd,
                               const double* inflow_vel, double scaling_factor, int incremental,
                               int proj_for_small_dt, double rtol, double atol, orc_stats* st)
{
    const int nx = p->n[0], ny = p->n[1], nz = p->n[2];
    const long ncell = (long)nx * ny * nz;
    const long sx = nx + 2 * ng, sy = ny + 2 * ng, sz = nz + 2 * ng;
    const long dsx = nx + 2 * ngd, dsy = ny + 2 * ngd;
    const long nnode = (long)(nx + 1) * (ny + 1) * (nz + 1);
#define VIDX(c, i, j, k) ((((long)(c) * sz + ((k) + ng)) * sy + ((j) + ng)) * sx + ((i) + ng))
#define DIDX(i, j, k) ((((long)((k) + ngd)) * dsy + ((j) + ngd)) * dsx + ((i) + ngd))
    /* :39-62  u += gp * s / rho on valid cells */
    if (!incremental) {
        for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
            double rho = density ? density[DIDX(i, j, k)] : ro_0;
            double soverrho = scaling_factor / rho;
            for (int c = 0; c < 3; ++c) velocity[VIDX(c, i, j, k)] += gp[c * ncell + ((long)k * ny + j) * nx + i] * soverrho;
        }
    }
    /* :65-71 */
    if (proj_for_small_dt || incremental)
        for (int c = 0; c < 3; ++c) for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i)
            velocity[VIDX(c, i, j, k)] -= velocity_o[VIDX(c, i, j, k)];
    int set_inflow_bc = !proj_for_small_dt && !incremental; /* :81 */
    /* :101-121 sigma = s / rho */
    double* sigma = NULL;
    if (density) {
        sigma = (double*)malloc(sizeof(double) * ncell);
        for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i)
            sigma[((long)k * ny + j) * nx + i] = scaling_factor / density[DIDX(i, j, k)];
    }
    /* :137 vel.setBndry(0) -- all ghost cells */
    for (int c = 0; c < 3; ++c) for (int k = -ng; k < nz + ng; ++k) for (int j = -ng; j < ny + ng; ++j) for (int i = -ng; i < nx + ng; ++i)
        if (i < 0 || i >= nx || j < 0 || j >= ny || k < 0 || k >= nz) velocity[VIDX(c, i, j, k)] = 0.0;
    /* :138-163 inflow faces: first ghost layer gets the Dirichlet inflow value */
    if (set_inflow_bc && inflow_vel) {
        for (int c = 0; c < 3; ++c) for (int k = -1; k <= nz; ++k) for (int j = -1; j <= ny; ++j) for (int i = -1; i <= nx; ++i) {
            int out = 0, inflow = 0;
            const int idx[3] = {i, j, k};
            for (int d = 0; d < 3; ++d) {
                if (idx[d] < 0) { out = 1; if (p->bclo[d] == ORC_BC_INFLOW) inflow = 1; }
                if (idx[d] >= p->n[d]) { out = 1; if (p->bchi[d] == ORC_BC_INFLOW) inflow = 1; }
            }
            if (out && inflow) velocity[VIDX(c, i, j, k)] = inflow_vel[VIDX(c, i, j, k)];
        }
    }
    /* :181-215 projector */
    double* phi = (double*)malloc(sizeof(double) * nnode);
    double* gphi = (double*)malloc(sizeof(double) * 3 * ncell);
    int status = orc_project(p, velocity, ng, sigma, scaling_factor / ro_0, phi, gphi, NULL, rtol, atol, st);
    /* :221-256 copy-out */
    if (incremental) {
        for (long i = 0; i < 3 * ncell; ++i) gp[i] += gphi[i];
        for (long i = 0; i < nnode; ++i) p_nd[i] += phi[i];
    } else {
        memcpy(gp, gphi, sizeof(double) * 3 * ncell);
        memcpy(p_nd, phi, sizeof(double) * nnode);
    }
    /* :86-92 */
    if (proj_for_small_dt || incremental)
        for (int c = 0; c < 3; ++c) for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i)
            velocity[VIDX(c, i, j, k)] += velocity_o[VIDX(c, i, j, k)];
    free(phi); free(gphi); free(sigma);
    return status;
#undef VIDX
#undef DIDX
}
