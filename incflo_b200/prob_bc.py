"""Host-side (numpy) restatement of IncfloVelFill, src/prob/prob_bc.H:8-351 (mass-inflow and direction_dependent
faces, probtypes 1101 / 1102), of HydroUtils::enforceInOutSolvability as called at
src/projection/incflo_apply_nodal_projection.cpp:166-179, and of incflo::make_nodalBC_mask
(src/boundary_conditions/incflo_set_bcs.cpp:10-53).

Used by the tests to build the `inflow_vel` array the oracle / the C ABI accept, and as the readable
statement of what the device kernel k_incflo_vel_fill (csrc/np_kernels.cuh) evaluates when a profile
has been set with b200np_set_inflow_profile.  bcv_vel[o][c]: o = amrex::Orientation index
(x-lo, y-lo, z-lo, x-hi, y-hi, z-hi), c = velocity component (incflo's m_bc_velocity)."""
import numpy as np

BC_INFLOW = 3
FACE_DEFAULT, FACE_DIRECTION_DEPENDENT, FACE_MIXED = 0, 1, 2


def _parab6(idx, n):
    s = (idx + 0.5) / n
    return 6.0 * s * (1.0 - s)


def _norm_profile(d, side, probtype, b, i, j, k, n, time):
    nx, ny, nz = n
    norm = np.full(i.shape, b[d])
    if d == 0 and side == 0:
        if probtype == 42: norm = np.full(i.shape, float(time))
        elif probtype == 31: norm = _parab6(j, ny)
        elif probtype == 43: norm = _parab6(j, ny) - 1.0
        elif probtype == 311: norm = _parab6(k, nz)
        elif probtype == 41: norm = 0.5 * ((k + 0.5) / nz)
    elif d == 0 and side == 1:
        if probtype == 42: norm = np.full(i.shape, float(time))
        elif probtype == 43: norm = _parab6(j, ny) - 1.0
    elif d == 1 and side == 0:
        if probtype == 32: norm = norm * _parab6(k, nz)
        if probtype == 322: norm = norm * _parab6(i, nx)
    elif d == 1 and side == 1:
        if probtype == 16:
            x = (i + 0.5) / nx
            norm = 16.0 * (x ** 4 - 2.0 * x ** 3 + x ** 2)
    elif d == 2 and side == 0:
        if probtype == 33: norm = norm * _parab6(i, nx)
        elif probtype == 333: norm = norm * _parab6(j, ny)
    return np.broadcast_to(norm, i.shape)


def incflo_vel_fill(n_cell, ng, bclo, bchi, probtype, bcv_vel, time=0.0, face_type=None, vel=None):
    """array shaped like the velocity, (3, nz+2ng, ny+2ng, nx+2ng), whose first ghost layer holds the values
    IncfloVelFill would write (zero elsewhere).  Faces are applied in the reference's order (x-lo, x-hi, y-lo, y-hi,
    z-lo, z-hi): a ghost cell outside in two directions keeps the later face's value.
    face_type (Orientation order, FACE_*): direction_dependent faces take the boundary value where the profile points
    into the domain and copy the first interior cell of `vel` (the velocity after setBndry(0)) where it points out --
    prob_bc.H:93-109 and siblings; the z-lo block's inflow test is norm_vel <= 0 as written at :300-301."""
    nx, ny, nz = n_cell
    src = np.zeros((3, nz + 2 * ng, ny + 2 * ng, nx + 2 * ng)) if vel is None else np.array(vel, dtype=np.float64, copy=True)
    # setBndry(0): everything outside the valid cells
    inner = src[:, ng:ng + nz, ng:ng + ny, ng:ng + nx].copy()
    src[:] = 0.0
    src[:, ng:ng + nz, ng:ng + ny, ng:ng + nx] = inner
    k, j, i = np.meshgrid(np.arange(-1, nz + 1), np.arange(-1, ny + 1), np.arange(-1, nx + 1), indexing="ij")
    idx = (i, j, k)
    n = (nx, ny, nz)
    view = src[:, ng - 1:ng + nz + 1, ng - 1:ng + ny + 1, ng - 1:ng + nx + 1]
    bcv = np.asarray(bcv_vel, dtype=np.float64).reshape(6, 3)
    ft = [FACE_DEFAULT] * 6 if face_type is None else list(face_type)
    for d in range(3):
        for side in range(2):
            bc = bclo[d] if side == 0 else bchi[d]
            if bc == 0:
                continue
            outside = idx[d] < 0 if side == 0 else idx[d] >= n[d]
            o = d + 3 * side
            b = bcv[o]
            if probtype == 1101 and d == 0:     # :86-92 / :140-146
                half = ny // 2
                sel = outside & ((j > half) if side == 0 else (j <= half))
                for c in range(3):
                    view[c][sel] = bcv[0][c] if side == 0 else -bcv[3][c]
                continue
            if probtype == 1102 and d == 1 and side == 1:   # :243-251
                sel = outside & (k <= nz // 2)
                for c in range(3):
                    view[c][sel] = -bcv[4][c]
                continue
            if bc != BC_INFLOW or ft[o] == FACE_MIXED:
                continue
            norm = _norm_profile(d, side, probtype, b, i, j, k, n, time)
            if ft[o] == FACE_DIRECTION_DEPENDENT:
                inflow = (norm >= 0.0) if (side == 0 and d != 2) else (norm <= 0.0)
            else:
                inflow = np.ones(i.shape, bool)
            ax = {0: 3, 1: 2, 2: 1}[d]
            shift = np.roll(view, -1 if side == 0 else 1, axis=ax)   # neighbour one cell towards the interior
            for c in range(3):
                val = norm if c == d else np.full(i.shape, b[c])
                sel_in = outside & inflow
                sel_out = outside & ~inflow
                newc = view[c].copy()
                newc[sel_in] = val[sel_in]
                newc[sel_out] = shift[c][sel_out]
                view[c][...] = newc
    out = src
    out[:, ng:ng + nz, ng:ng + ny, ng:ng + nx] = 0.0
    return out


def enforce_inout_solvability(vel, n_cell, ng, dx, face_type, small_vel=1.0e-8):
    """HydroUtils::enforceInOutSolvability on the first ghost layer of the direction_dependent faces (face cells only,
    no edge / corner ghost cells): outflow values *= influx / outflux.  Returns (influx, outflux); raises where
    AMReX-Hydro aborts.  vel is modified in place."""
    nx, ny, nz = n_cell
    n = (nx, ny, nz)
    influx = outflux = 0.0
    faces = []
    for o in range(6):
        if face_type[o] != FACE_DIRECTION_DEPENDENT:
            continue
        d, side = o % 3, o // 3
        sl = [d, slice(ng, ng + nz), slice(ng, ng + ny), slice(ng, ng + nx)]
        sl[{0: 3, 1: 2, 2: 1}[d]] = ng - 1 if side == 0 else ng + n[d]
        v = vel[tuple(sl)]
        ds = dx[(d + 1) % 3] * dx[(d + 2) % 3]
        inflow = (v >= 0.0) if side == 0 else (v <= 0.0)
        influx += np.abs(v[inflow]).sum() * ds
        outflux += np.abs(v[~inflow]).sum() * ds
        faces.append((tuple(sl), inflow))
    if (influx > small_vel) != (outflux > small_vel):
        raise RuntimeError("Cannot enforce solvability: inflow without outflow through the direction_dependent faces, or the reverse")
    if influx > small_vel:
        alpha = influx / outflux
        for sl, inflow in faces:
            v = vel[sl]
            v[~inflow] *= alpha
            vel[sl] = v
    return influx, outflux


def make_nodalBC_mask(n_cell, face_type, split_dir, half_num_cells):
    """incflo::make_nodalBC_mask + prob_set_BC_MF (probtypes 1100/1101/1102): nodal int32 mask (nz+1, ny+1, nx+1), 1 = solve,
    0 = outflow (Dirichlet) part of a mixed face: low side idx[split_dir] <= half, high side idx[split_dir] > half"""
    nx, ny, nz = n_cell
    mask = np.ones((nz + 1, ny + 1, nx + 1), dtype=np.int32)
    k, j, i = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    idx = (i, j, k)
    n = (nx, ny, nz)
    for o in range(6):
        if face_type[o] != FACE_MIXED:
            continue
        d, side = o % 3, o // 3
        on_face = idx[d] == (0 if side == 0 else n[d])
        out = (idx[split_dir] <= half_num_cells) if side == 0 else (idx[split_dir] > half_num_cells)
        mask[on_face & out] = 0
    return mask
