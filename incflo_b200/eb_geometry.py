"""Synthetic embedded-boundary geometry for the EB nodal projection (BASELINE configs[4], test_3d/benchmark.channel_cylinder-x).

Stand-in for what the caller owns in the reference: AMReX EB2 builds the geometry from an implicit function
(src/embedded_boundaries/eb_cylinder.cpp:15-64: EB2::CylinderIF -> GeometryShop -> EB2::Build) and
amrex::EBFArrayBoxFactory hands out, per cell, the volume fraction, the boundary normal / area / centroid; from those
MLNodeLaplacian::buildIntegral computes 18 monomial integrals over the fluid part of the cell.  EB
geometry generation is OUT OF SCOPE of the projection path; this module only produces inputs of the
same kind for tests, tools and bench: like EB2 it replaces the body inside a cut cell by ONE half-space and then
integrates monomials over (unit cell) n (half-space) exactly (tetrahedra + Gauss quadrature of sufficient degree).

Conventions (AMReX): cell-local coordinates x, y, z in [-1/2, 1/2] with the origin at the cell centre, lengths in
units of the (isotropic) cell size; the implicit function is negative in the fluid; the boundary normal points out of the
fluid into the body ("The EB normal points out of the domain", src/boundary_conditions/incflo_set_bcs.cpp:265-267).
Monomial order of the volume integrals = amrex i_S_* (AMReX_MLNodeLap_K.H, restated from memory [U]):
    S_x, S_y, S_z, S_x2, S_y2, S_z2, S_x_y, S_x_z, S_y_z, S_x2_y, S_x2_z, S_x_y2, S_y2_z, S_x_z2, S_y_z2, S_x2_y2, S_x2_z2, S_y2_z2
Surface integrals over the EB face inside the cell (i_B_*): B_1 (= boundary area), B_x, B_y, B_z, B_x_y, B_x_z, B_y_z, B_xyz.
Arrays: (ncomp, nz, ny, nx), C order (i fastest, component outermost = amrex::Array4).
"""
import itertools

import numpy as np

# exponents (px, py, pz) of the 18 volume integrals, AMReX order
S_EXP = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (2, 0, 0), (0, 2, 0), (0, 0, 2), (1, 1, 0), (1, 0, 1), (0, 1, 1),
         (2, 1, 0), (2, 0, 1), (1, 2, 0), (0, 2, 1), (1, 0, 2), (0, 1, 2), (2, 2, 0), (2, 0, 2), (0, 2, 2)]
B_EXP = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1)]
N_SINTG, N_BINTG = 18, 8

_CORNERS = np.array(list(itertools.product((-0.5, 0.5), repeat=3)))              # (8, 3): x slowest here, irrelevant
_EDGES = [(a, b) for a in range(8) for b in range(a + 1, 8) if np.sum(_CORNERS[a] != _CORNERS[b]) == 1]
_GX, _GW = np.polynomial.legendre.leggauss(4)
_GX, _GW = 0.5 * (_GX + 1.0), 0.5 * _GW                                           # on [0, 1], exact to degree 7


def regular_integrals():
    """the 18 integrals of an uncut cell: int x^2 = 1/12, int x^2 y^2 = 1/144, odd ones vanish"""
    out = np.zeros(N_SINTG)
    for m, (px, py, pz) in enumerate(S_EXP):
        v = 1.0
        for p in (px, py, pz):
            v *= 0.0 if p == 1 else (1.0 / 12.0 if p == 2 else 1.0)
        out[m] = v
    return out


def _tet_monomials(v, exps):
    """integrals of x^px y^py z^pz over the tetrahedron with vertices v (4, 3): collapsed coordinates + 4-point Gauss per
    direction (exact for total degree <= 4)"""
    r, s, t = np.meshgrid(_GX, _GX, _GX, indexing="ij")
    w = _GW[:, None, None] * _GW[None, :, None] * _GW[None, None, :] * (1 - r) ** 2 * (1 - s)
    l1, l2, l3 = r, s * (1 - r), t * (1 - r) * (1 - s)
    l0 = 1 - l1 - l2 - l3
    p = l0[..., None] * v[0] + l1[..., None] * v[1] + l2[..., None] * v[2] + l3[..., None] * v[3]
    vol6 = abs(np.linalg.det(v[1:] - v[0]))
    return np.array([np.sum(w * p[..., 0] ** a * p[..., 1] ** b * p[..., 2] ** c) for a, b, c in exps]) * vol6


def _tri_monomials(v, exps):
    """integrals over the triangle v (3, 3) (surface measure), exact for degree <= 3 and beyond"""
    r, s = np.meshgrid(_GX, _GX, indexing="ij")
    w = _GW[:, None] * _GW[None, :] * (1 - r)
    l1, l2 = r, s * (1 - r)
    l0 = 1 - l1 - l2
    p = l0[..., None] * v[0] + l1[..., None] * v[1] + l2[..., None] * v[2]
    area2 = np.linalg.norm(np.cross(v[1] - v[0], v[2] - v[0]))
    return np.array([np.sum(w * p[..., 0] ** a * p[..., 1] ** b * p[..., 2] ** c) for a, b, c in exps]) * area2


def cut_cell_simplices(normal, offset):
    """(tets, tris) of the fluid part of the unit cell cut by normal . x = offset: tets (nt, 4, 3) fill the fluid polyhedron,
    tris (ns, 3, 3) the part of the plane inside the cell.  None for an uncut cell.  (Used by the golden generator to integrate
    shape-function products directly, without going through the monomial integrals.)"""
    from scipy.spatial import ConvexHull
    n = np.asarray(normal, dtype=np.float64)
    n = n / np.linalg.norm(n)
    sd = _CORNERS @ n - offset
    tol = 1e-13
    if np.all(sd <= tol) or np.all(sd >= -tol):
        return None
    pts = [c for c, s in zip(_CORNERS, sd) if s <= tol]
    for a, b in _EDGES:
        if (sd[a] < -tol and sd[b] > tol) or (sd[a] > tol and sd[b] < -tol):
            pts.append(_CORNERS[a] + sd[a] / (sd[a] - sd[b]) * (_CORNERS[b] - _CORNERS[a]))
    pts = np.array(pts)
    hull = ConvexHull(pts)
    c = pts[hull.vertices].mean(axis=0)
    tets = np.array([np.vstack([c[None], pts[t]]) for t in hull.simplices])
    tris = [pts[t] for t in hull.simplices if np.all(np.abs(pts[t] @ n - offset) < 1e-11)]
    return tets, np.array(tris)


def cut_cell(normal, offset):
    """Unit cell [-1/2, 1/2]^3 cut by the plane normal . x = offset; the fluid is the side normal . x <= offset (the normal
    points into the body).  Returns (vfrac, S[18], B[8]) with S the volume integrals over the fluid part and B the surface
    integrals over the part of the plane inside the cell.  Fully fluid / fully covered cells come out as V = 1 / 0, B = 0."""
    from scipy.spatial import ConvexHull
    n = np.asarray(normal, dtype=np.float64)
    n = n / np.linalg.norm(n)
    sd = _CORNERS @ n - offset                                  # signed distance of the corners: <= 0 fluid
    tol = 1e-13
    if np.all(sd <= tol):
        return 1.0, regular_integrals(), np.zeros(N_BINTG)
    if np.all(sd >= -tol):
        return 0.0, np.zeros(N_SINTG), np.zeros(N_BINTG)
    pts = [c for c, s in zip(_CORNERS, sd) if s <= tol]
    onplane = []
    for a, b in _EDGES:
        if (sd[a] < -tol and sd[b] > tol) or (sd[a] > tol and sd[b] < -tol):
            t = sd[a] / (sd[a] - sd[b])
            q = _CORNERS[a] + t * (_CORNERS[b] - _CORNERS[a])
            pts.append(q)
            onplane.append(q)
    pts = np.array(pts)
    if len(pts) < 4:
        return 0.0, np.zeros(N_SINTG), np.zeros(N_BINTG)
    try:
        hull = ConvexHull(pts)
    except Exception:                                            # flat sliver: no volume
        return 0.0, np.zeros(N_SINTG), np.zeros(N_BINTG)
    c = pts[hull.vertices].mean(axis=0)
    S = np.zeros(1 + N_SINTG)
    B = np.zeros(N_BINTG)
    for tri in hull.simplices:
        v = pts[tri]
        S += _tet_monomials(np.vstack([c[None], v]), [(0, 0, 0)] + S_EXP)
        if np.all(np.abs(v @ n - offset) < 1e-11):              # this facet lies on the cutting plane
            B += _tri_monomials(v, B_EXP)
    return S[0], S[1:], B


class EBGeometry:
    """per-cell EB data of one level: vfrac (nz, ny, nx), intg (18, ...), bnorm (3, ...), barea (...), bintg (8, ...)"""

    def __init__(self, n):
        nx, ny, nz = n
        self.n = tuple(n)
        self.vfrac = np.ones((nz, ny, nx))
        self.intg = np.empty((N_SINTG, nz, ny, nx))
        self.intg[:] = regular_integrals()[:, None, None, None]
        self.bnorm = np.zeros((3, nz, ny, nx))
        self.bintg = np.zeros((N_BINTG, nz, ny, nx))
        self.boff = np.zeros((nz, ny, nx))                      # plane offset of a cut cell: the fluid is bnorm . x <= boff

    @property
    def barea(self):
        return self.bintg[0]

    def cut_mask(self):
        """EBCellFlag::isSingleValued"""
        return (self.vfrac > 0.0) & (self.vfrac < 1.0)

    def set_cell(self, i, j, k, V, S, B, normal, off=0.0):
        self.vfrac[k, j, i] = V
        self.boff[k, j, i] = off if 0.0 < V < 1.0 else 0.0
        self.intg[:, k, j, i] = S
        self.bintg[:, k, j, i] = B
        self.bnorm[:, k, j, i] = normal if 0.0 < V < 1.0 else 0.0


def from_implicit(n, h, f, gradf, prob_lo=(0.0, 0.0, 0.0), extruded_z=False, small_vfrac=1e-9):
    """f(x, y, z) < 0 in the fluid (amrex EB2 convention), gradf its gradient.  Every cell whose 8 corners do not agree in
    sign is cut by the plane through the point of the zero level set closest (one Newton step along gradf) to the cell centre,
    with the normal gradf / |gradf| there.  Cut cells with vfrac < small_vfrac become covered, > 1 - small_vfrac regular
    (EB2 does the same, eb2.small_volfrac).  The monomial-integral form of the operator loses digits in tiny cells by cancellation
    (int (1/2 - y)^2 = V/4 - S_y + S_y2 ...): with vfrac ~ 1e-5 its entries are good to ~1e-7 only -- in AMReX as well."""
    nx, ny, nz = n
    g = EBGeometry(n)
    xs = prob_lo[0] + h * np.arange(nx + 1)
    ys = prob_lo[1] + h * np.arange(ny + 1)
    zs = prob_lo[2] + h * np.arange(nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    F = f(X, Y, Z)
    inside = F >= 0.0                                            # node in the body
    cnt = np.zeros((nz, ny, nx), dtype=np.int32)
    for dk, dj, di in itertools.product((0, 1), repeat=3):
        cnt += inside[dk:dk + nz, dj:dj + ny, di:di + nx]
    covered = cnt == 8
    g.vfrac[covered] = 0.0
    g.intg[:, covered] = 0.0
    cache = {}
    ks, js, is_ = np.nonzero((cnt > 0) & (cnt < 8))
    for k, j, i in zip(ks, js, is_):
        key = (j, i) if extruded_z else (k, j, i)
        if key not in cache:
            c = np.array([xs[i] + 0.5 * h, ys[j] + 0.5 * h, zs[k] + 0.5 * h])
            p = c.copy()
            for _ in range(8):                                   # closest point of f = 0 along the gradient
                gr = np.asarray(gradf(*p), dtype=np.float64)
                p = p - f(*p) * gr / np.dot(gr, gr)
            gr = np.asarray(gradf(*p), dtype=np.float64)
            nrm = gr / np.linalg.norm(gr)                        # f grows into the body: the normal points into the body
            off = np.dot(nrm, (p - c) / h)
            V, S, B = cut_cell(nrm, off)
            if V < small_vfrac:
                V, S, B = 0.0, np.zeros(N_SINTG), np.zeros(N_BINTG)
            elif V > 1.0 - small_vfrac:
                V, S, B = 1.0, regular_integrals(), np.zeros(N_BINTG)
            cache[key] = (V, S, B, nrm, off)
        g.set_cell(i, j, k, *cache[key])
    return g


def cylinder(n, h, radius, center, direction=2, internal_flow=False, prob_lo=(0.0, 0.0, 0.0), small_vfrac=1e-9):
    """EB2::CylinderIF(radius, direction, center, inside) of make_eb_cylinder (src/embedded_boundaries/eb_cylinder.cpp:44):
    an infinite cylinder along `direction`; internal_flow = False: the fluid is outside (channel_cylinder-x)."""
    ax = [d for d in range(3) if d != direction]
    sgn = -1.0 if internal_flow else 1.0

    def f(x, y, z):
        q = (x, y, z)
        r = np.sqrt((q[ax[0]] - center[ax[0]]) ** 2 + (q[ax[1]] - center[ax[1]]) ** 2)
        return sgn * (radius - r)

    def gradf(x, y, z):
        q = (x, y, z)
        r = np.sqrt((q[ax[0]] - center[ax[0]]) ** 2 + (q[ax[1]] - center[ax[1]]) ** 2)
        gr = np.zeros(3)
        gr[ax[0]] = -sgn * (q[ax[0]] - center[ax[0]]) / r
        gr[ax[1]] = -sgn * (q[ax[1]] - center[ax[1]]) / r
        return gr

    return from_implicit(n, h, f, gradf, prob_lo, extruded_z=(direction == 2), small_vfrac=small_vfrac)


def sphere(n, h, radius, center, internal_flow=False, prob_lo=(0.0, 0.0, 0.0), small_vfrac=1e-9):
    """EB2::SphereIF: test_3d/benchmark.uniform_velocity_sphere-type geometry"""
    sgn = -1.0 if internal_flow else 1.0
    c = np.asarray(center, dtype=np.float64)

    def f(x, y, z):
        return sgn * (radius - np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2))

    def gradf(x, y, z):
        d = np.array([x - c[0], y - c[1], z - c[2]])
        return -sgn * d / np.linalg.norm(d)

    return from_implicit(n, h, f, gradf, prob_lo, small_vfrac=small_vfrac)


def plane(n, h, point, normal, prob_lo=(0.0, 0.0, 0.0), small_vfrac=1e-9):
    """EB2::PlaneIF: body on the side the normal points to"""
    nrm = np.asarray(normal, dtype=np.float64)
    nrm = nrm / np.linalg.norm(nrm)
    pt = np.asarray(point, dtype=np.float64)
    return from_implicit(n, h, lambda x, y, z: (x - pt[0]) * nrm[0] + (y - pt[1]) * nrm[1] + (z - pt[2]) * nrm[2],
                         lambda x, y, z: nrm, prob_lo, small_vfrac=small_vfrac)
