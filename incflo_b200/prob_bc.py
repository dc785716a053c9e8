"""Host-side (numpy) restatement of IncfloVelFill, src/prob/prob_bc.H:8-351, mass-inflow (ext_dir) faces.

Used by the tests to build the `inflow_vel` array the oracle / the C ABI accept, and as the readable
statement of what the device kernel k_incflo_vel_fill (csrc/np_kernels.cuh) evaluates when a profile
has been set with b200np_set_inflow_profile.  bcv_vel[o][c]: o = amrex::Orientation index
(x-lo, y-lo, z-lo, x-hi, y-hi, z-hi), c = velocity component (incflo's m_bc_velocity)."""
import numpy as np

BC_INFLOW = 3


def _parab6(idx, n):
    s = (idx + 0.5) / n
    return 6.0 * s * (1.0 - s)


def incflo_vel_fill(n_cell, ng, bclo, bchi, probtype, bcv_vel, time=0.0):
    """array shaped like the velocity, (3, nz+2ng, ny+2ng, nx+2ng), whose first ghost layer at INFLOW faces
    holds the values IncfloVelFill would write (zero elsewhere).  Faces are applied in the reference's order
    (x-lo, x-hi, y-lo, y-hi, z-lo, z-hi): a ghost cell outside in two directions keeps the later face's value."""
    nx, ny, nz = n_cell
    out = np.zeros((3, nz + 2 * ng, ny + 2 * ng, nx + 2 * ng))
    k, j, i = np.meshgrid(np.arange(-1, nz + 1), np.arange(-1, ny + 1), np.arange(-1, nx + 1), indexing="ij")
    idx = (i, j, k)
    n = (nx, ny, nz)
    view = out[:, ng - 1:ng + nz + 1, ng - 1:ng + ny + 1, ng - 1:ng + nx + 1]
    bcv = np.asarray(bcv_vel, dtype=np.float64).reshape(6, 3)
    for d in range(3):
        for side in range(2):
            if (bclo[d] if side == 0 else bchi[d]) != BC_INFLOW:
                continue
            outside = idx[d] < 0 if side == 0 else idx[d] >= n[d]
            b = bcv[d + 3 * side]
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
            for c in range(3):
                val = norm if c == d else np.full(i.shape, b[c])
                view[c][outside] = np.broadcast_to(val, i.shape)[outside]
    return out
