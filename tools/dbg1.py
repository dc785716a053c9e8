import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from helpers import BC_CASES, oracle_params, random_sigma, TILE
from oracle import pyoracle as po
from incflo_b200 import nodal_projector as npj
case = BC_CASES[2]
name, n, dx, bclo, bchi = case
rng = np.random.default_rng(0)
sigma = random_sigma(n, rng)
p = oracle_params(n, dx, bclo, bchi)
mg = po.MG(p, sigma, 0.7)
vel = np.zeros((3, n[2] + 2, n[1] + 2, n[0] + 2))
proj = npj.NodalProjector(vel, sigma, 0.7, dict(n_cell=n, dx=dx, is_periodic=[b == 0 for b in bclo]), ng=1)
proj.setDomainBC(bclo, bchi)
proj.set_sigma(sigma, 0.7)
for lev in range(mg.nlev):
    gs = proj.level_get(lev, npj.A_SIGMA); os_ = mg.sigma(lev)
    print("lev", lev, mg.dims(lev), "sigma err", np.abs(gs-os_).max())
    phi = rng.standard_normal(mg.node_shape(lev)); rhs = rng.standard_normal(mg.node_shape(lev))
    proj.level_set(lev, npj.A_COR, phi); proj.level_set(lev, npj.A_RES, rhs)
    proj.level_op(lev, npj.OP_RESIDUAL)
    got = proj.level_get(lev, npj.A_RESCOR)
    ref = mg.residual(lev, phi, rhs)
    err = np.abs(got-ref)
    print("   resid err max", err.max()/np.abs(ref).max(), "at", np.unravel_index(err.argmax(), err.shape), "n bad", (err > 1e-10*np.abs(ref).max()).sum())
    bad = np.argwhere(err > 1e-10*np.abs(ref).max())
    print("   bad k:", sorted(set(bad[:,0])), "j:", sorted(set(bad[:,1])), "i:", sorted(set(bad[:,2])))
