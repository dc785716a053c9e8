#!/usr/bin/env python
"""bench.py -- nodal-projection throughput (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n 256]

One "step" = one incflo::ApplyNodalProjection-equivalent on the workload BASELINE.json quotes the
metric on at one GPU: configs[1], test_no_eb_3d/benchmark.rayleigh_taylor, variable density
(sigma = dt/rho, 4:1 contrast), 256^3, periodic x/y + slip walls in z, solved to
nodal_proj.mg_rtol = 1e-11 / mg_atol = 1e-14 with the reference's default V(2,2) x 4-sweep
cycle: pre-add u += dt gp/rho, sigma build, rhs = D u, MLMG solve, u -= sigma G phi,
gp = G phi, p = phi.  Synthetic closed-form fields (incflo_b200/problems.py).

value  : Mcell-updates/s = cells * K / t with every input already resident in HBM (device
         pointers through the C ABI), timed with CUDA events on the launching stream.
e2e    : same metric through the same C-ABI call with HOST (pinned) buffers: H2D of
         velocity/density/gp and D2H of velocity/gp/p_nd inside the timed region.
roofline: the dominant kernel (tile-resident Gauss-Seidel sweep, level 0): algorithmic bytes
         (32 B/node variable sigma) / CUDA-event time per launch vs MEASURED_PEAKS.json hbm_gbs
         (burst copy figure; the kernel is timed alone, back to back); traffic = dram bytes of one
         launch from the committed ncu --set full capture (profiles/ncu_traffic.json).
parity : before anything is timed, the same C-ABI call is checked against the CPU oracle (mirrored
         smoother ordering) on a size whose level 0 runs the PRODUCTION kernels: N = 1: 128^3;
         N > 1: 256 x 256 x 32N cells on N ranks (level 0 = the fused NVLink-halo sweep).  rel-L2 of
         p, u, gp must be < 1e-9 (north_star), else the run exits non-zero.
cpu_baseline / --impl reference: the CPU oracle configured as the REFERENCE's CPU algorithm
         (AMReX multi-box semantics, SURVEY A.4: lexicographic Gauss-Seidel inside each
         max_grid_size box, 4 sweeps per smooth call without halo refresh) on the FULL 256^3
         workload with OpenMP on all host cores (the count in effect is printed).  It is a port:
         incflo/AMReX itself cannot be built offline (AMReX, AMReX-Hydro, MPI are not vendored).
N > 1  : z-slab decomposition, one process per GPU (torchrun).  Headline = weak scaling: every rank
         owns 256 x 256 x 256 cells of a 256 x 256 x 256N domain.  The line also carries a `strong`
         record: 512^3 and 1024^3 in total on the N GPUs against the same solve MEASURED on one GPU
         (rank 0) in the same run.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "nodal_projection_Mcell_updates_per_s"
UNIT = "Mcell-updates/s"
RTOL, ATOL = 1e-11, 1e-14
PARITY_TOL = 1e-9            # north_star: pressure and projected velocity within 1e-9 relative L2
REF_BOX = 64                 # amr.max_grid_size of the CPU reference arm at 256^3 (64 boxes)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def solve_roofline(nodes, cells, vcycles, var_sigma=True):
    """algorithmic HBM bytes of one projection from SURVEY 8(d)'s per-kernel table: one V(2,2) x 4-sweep cycle is
    16 sweeps + residual + restriction + interpolation + (sol += cor, top residual) = 634 B per fine node (variable
    sigma; 490 B constant), x 8/7 for the hierarchy, plus rhs (32 B/node) and pre-add + final update (88 + 96 B/cell)"""
    per_node_cycle = (16 * 32 + 32 + 9 + 25 + 56) if var_sigma else (16 * 24 + 24 + 9 + 17 + 48)
    per_cell = (88 + 96) if var_sigma else (72 + 88)
    return nodes * (per_node_cycle * 8.0 / 7.0 * vcycles + 32.0) + cells * float(per_cell)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# workloads (closed forms, filled in z chunks so that a 1024^3 slab never needs a second copy)
# ------------------------------------------------------------------------------------------
def fill_inputs(vel, gp, n_glob, zlo, nz, ng, rho=None, chunk=64):
    """(re)generate this rank's inputs in place: velocity (valid cells; ghost cells zero), gp, and optionally rho"""
    from incflo_b200 import problems
    N0, N1 = n_glob[0], n_glob[1]
    vel.zero_()
    for z0 in range(0, nz, chunk):
        c = min(chunk, nz - z0)
        with problems.slab(zlo + z0, c):   # cell planes [zlo+z0, zlo+z0+c) of the global closed forms
            v = problems.rayleigh_taylor_velocity(n_glob, 0, vel.device, "b")
            vel[:, ng + z0:ng + z0 + c, ng:ng + N1, ng:ng + N0] = v
            if rho is not None:
                rho[ng + z0:ng + z0 + c, ng:ng + N1, ng:ng + N0] = problems.rayleigh_taylor_density(n_glob, 0, vel.device)
        del v
    gp.zero_()
    gp[2] = -0.05  # a hydrostatic-like old pressure gradient so the pre-add does work


def workload(n_glob, nranks, rank, device, ng=3):
    """this rank's z slab of the rayleigh_taylor workload on the global domain n_glob (dx = 1/n_glob[0], isotropic)"""
    import torch
    N0, N1, N2 = n_glob
    nz = N2 // nranks
    zlo = rank * nz
    vel = torch.empty((3, nz + 2 * ng, N1 + 2 * ng, N0 + 2 * ng), dtype=torch.float64, device=device)
    rho = torch.ones((nz + 2 * ng, N1 + 2 * ng, N0 + 2 * ng), dtype=torch.float64, device=device)
    gp = torch.empty((3, nz, N1, N0), dtype=torch.float64, device=device)
    p = torch.zeros((nz + 1, N1 + 1, N0 + 1), dtype=torch.float64, device=device)
    fill_inputs(vel, gp, n_glob, zlo, nz, ng, rho)
    return dict(n=tuple(n_glob), dx=(1.0 / N0,) * 3, dt=0.45 / N0, vel=vel, rho=rho, gp=gp, p=p, ng=ng, zlo=zlo, nz=nz,
                bclo=(0, 0, 1), bchi=(0, 0, 1))


# ------------------------------------------------------------------------------------------
# CPU oracle legs (checker for `parity`, CPU baseline; never part of the product path)
# ------------------------------------------------------------------------------------------
def oracle_threads():
    """torchrun exports OMP_NUM_THREADS=1: set the team size explicitly, report what is in effect"""
    from oracle import pyoracle as po
    return po.set_num_threads(os.cpu_count() or 1)


def oracle_params(n, mode):
    from oracle import pyoracle as po
    dx = (1.0 / n[0],) * 3
    if mode == "mirror":      # the GPU's tile ordering: sweep-by-sweep comparable
        return po.make_params(n, dx, (0, 0, 1), (0, 0, 1), smoother=po.SM_BOX, box=(64, 16, 64), box_order=po.SM_PLANE4,
                              box_stale_per_call=0)
    b = min(REF_BOX, n[0] // 2)   # the reference's CPU algorithm: multi-box lexicographic GS, 4 stale sweeps
    return po.make_params(n, dx, (0, 0, 1), (0, 0, 1), smoother=po.SM_BOX, box=(b, b, b), box_order=po.SM_LEX,
                          box_stale_per_call=1, box_amrex=1)


def cpu_port_run(n, steps, warmup, mode="reference", keep=False):
    """the CPU oracle on the rayleigh_taylor workload of global size n; returns (seconds per step, V-cycles, fields)"""
    import numpy as np
    from oracle import pyoracle as po
    ng = 3
    wl = workload(n, 1, 0, "cpu", ng)
    vel0, rho, gp0 = wl["vel"].numpy(), wl["rho"].numpy(), wl["gp"].numpy()
    prm = oracle_params(n, mode)
    times, iters, out = [], 0, None
    for s in range(warmup + steps):
        vel = vel0.copy(); gp = gp0.copy(); p = np.zeros((n[2] + 1, n[1] + 1, n[0] + 1))
        t0 = time.perf_counter()
        status, st = po.apply_nodal_projection(prm, vel, ng, gp, p, density=rho, ngd=ng, scaling_factor=wl["dt"],
                                               rtol=RTOL, atol=ATOL)
        dt = time.perf_counter() - t0
        assert status == 0
        iters = st.iters
        if s >= warmup:
            times.append(dt)
        if keep:
            out = dict(vel=vel, gp=gp, p=p)
    return times, iters, out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = oracle_threads()
    n = (args.n,) * 3
    steps, warmup = max(1, min(args.steps, 2)), min(args.warmup, 1)
    times, iters, _ = cpu_port_run(n, steps, warmup, "reference")
    t = sum(times) / len(times)
    val = args.n ** 3 / t / 1e6
    what = (f"CPU restatement of the reference algorithm (AMReX multi-box semantics: lexicographic Gauss-Seidel inside "
            f"{min(REF_BOX, args.n // 2)}^3-cell boxes, 4 sweeps per smooth call without halo refresh, V(2,2), BiCGStab bottom), OpenMP")
    sample = (f"the full {args.n}^3 rayleigh_taylor variable-density projection per step"
              + (f" (= the per-GPU share of the {args.gpus}-GPU weak-scaling workload)" if args.gpus > 1 else "")
              + f", {len(times)} timed steps after {warmup} warm-up")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
            "warmup": warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"rayleigh_taylor variable-density (sigma=dt/rho, 4:1) {args.n}^3, periodic x/y + walls z, nodal "
                                   f"projection to rtol 1e-11 (BASELINE configs[1])", "rtol": RTOL, "atol": ATOL, "vcycles": iters,
                       "cycle": "V(2,2) x 4 sweeps (reference defaults)", "algorithm": what},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "host_cpus": os.cpu_count(),
                             "note": "port, not incflo/AMReX itself: AMReX, AMReX-Hydro and MPI are not vendored and cannot be built offline"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
class Ctx:
    pass


def make_projection(ctx, n_glob, nranks=None):
    """a handle over the global domain n_glob on all ranks (or on this rank alone when nranks == 1)"""
    import torch
    import torch.distributed as dist
    from incflo_b200 import nodal_projector as npj
    nranks = ctx.nranks if nranks is None else nranks
    nccl_id = None
    if nranks > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device=ctx.device)
        if ctx.rank == 0:
            idt.copy_(torch.tensor(list(npj.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().tolist())
    proj = npj.IncfloProjection(n_glob, (1.0 / n_glob[0],) * 3, (0, 0, 1), (0, 0, 1), device=ctx.local,
                                rank=ctx.rank if nranks > 1 else 0, nranks=nranks, nccl_id=nccl_id)
    proj.set_stream(ctx.stream.cuda_stream)
    return proj


def allmax(ctx, x):
    import torch
    import torch.distributed as dist
    if ctx.nranks == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=ctx.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def parity_block(ctx):
    """C-ABI call vs the CPU oracle (mirrored ordering) on a size whose level 0 runs the production kernels"""
    import numpy as np
    import torch
    import torch.distributed as dist
    P, rank = ctx.nranks, ctx.rank
    n = (128, 128, 128) if P == 1 else (256, 256, 32 * P)
    wl = workload(n, P, rank, ctx.device)
    proj = make_projection(ctx, n)
    with torch.cuda.stream(ctx.stream):
        st = proj.apply_nodal_projection(wl["vel"], wl["ng"], wl["gp"], wl["p"], density=wl["rho"], ngd=wl["ng"],
                                         scaling_factor=wl["dt"], mg_rtol=RTOL, mg_atol=ATOL)
    torch.cuda.synchronize()
    transport = proj.halo_transport()
    pmap = proj.peer_map()
    proj.close()
    ng, nz, zlo = wl["ng"], wl["nz"], wl["zlo"]
    # oracle on the GLOBAL problem (rank 0), results broadcast as device tensors
    shapes = dict(vel=(3, n[2] + 2 * ng, n[1] + 2 * ng, n[0] + 2 * ng), gp=(3, n[2], n[1], n[0]), p=(n[2] + 1, n[1] + 1, n[0] + 1))
    ref, info = {}, torch.zeros(3, dtype=torch.float64, device=ctx.device)
    if rank == 0:
        t0 = time.perf_counter()
        _, it_m, out = cpu_port_run(n, 1, 0, "mirror", keep=True)
        info[0] = it_m; info[1] = time.perf_counter() - t0
        if P == 1:   # the reference's CPU algorithm on the same input: V-cycle count for the 20 % band
            _, it_r, _ = cpu_port_run(n, 1, 0, "reference")
            info[2] = it_r
    for k in ("vel", "gp", "p"):
        ref[k] = torch.from_numpy(out[k]).to(ctx.device) if rank == 0 else torch.empty(shapes[k], dtype=torch.float64, device=ctx.device)
        if P > 1:
            dist.broadcast(ref[k], 0)
    if P > 1:
        dist.broadcast(info, 0)
    inner = (slice(None), slice(ng, ng + nz), slice(ng, -ng), slice(ng, -ng))
    rv = ref["vel"][:, zlo:zlo + nz + 2 * ng][inner]
    rg = ref["gp"][:, zlo:zlo + nz]
    own = nz + (1 if rank == P - 1 else 0)          # uniquely owned node planes
    rp = ref["p"][zlo:zlo + own]
    dp = wl["p"][:own] - rp
    sums = torch.stack([((wl["vel"][inner] - rv) ** 2).sum(), (rv ** 2).sum(), ((wl["gp"] - rg) ** 2).sum(), (rg ** 2).sum(),
                        dp.sum(), torch.tensor(float(dp.numel()), dtype=torch.float64, device=ctx.device), rp.sum()])
    if P > 1:
        dist.all_reduce(sums)
    mean_d, mean_r = sums[4] / sums[5], sums[6] / sums[5]   # the problem is singular: compare p up to a constant
    s2 = torch.stack([((dp - mean_d) ** 2).sum(), ((rp - mean_r) ** 2).sum()])
    if P > 1:
        dist.all_reduce(s2)
    res = {"n_cell": list(n), "ranks": P, "checker": "CPU oracle, mirrored smoother ordering (oracle/nodal_oracle.c)",
           "rel_l2_u": float(torch.sqrt(sums[0] / sums[1])), "rel_l2_gp": float(torch.sqrt(sums[2] / sums[3])),
           "rel_l2_p": float(torch.sqrt(s2[0] / s2[1])), "tol": PARITY_TOL, "vcycles_gpu": int(st.iters),
           "vcycles_oracle_mirrored": int(info[0].item()), "oracle_seconds": float(info[1].item()),
           "halo_transport": {0: "none (1 GPU)", 1: "peer memory", 2: "nccl"}[transport],
           "peer_map": {0: None, 1: "cuMem + POSIX fd", 2: "cudaIpc"}[pmap]}
    if P == 1:
        res["vcycles_reference_cpu_algorithm"] = int(info[2].item())
    res["ok"] = bool(st.status == 0 and max(res["rel_l2_u"], res["rel_l2_gp"], res["rel_l2_p"]) < PARITY_TOL
                     and abs(res["vcycles_gpu"] - res["vcycles_oracle_mirrored"]) <= 1)
    del wl, ref
    torch.cuda.empty_cache()
    return res


def timed_solves(ctx, proj, wl, K, W, n_glob, restore="clone"):
    """W warm-up + K timed steps; returns (ms per step (max over ranks), per-step list, last stats, launches)"""
    import torch
    import torch.distributed as dist
    ng = wl["ng"]

    def step(vel, gp, p):
        return proj.apply_nodal_projection(vel, ng, gp, p, density=wl["rho"], ngd=ng, scaling_factor=wl["dt"],
                                           mg_rtol=RTOL, mg_atol=ATOL)
    if restore == "clone":      # every timed step of a chunk gets its own untouched inputs
        nbuf = min(K, 16)
        vels = [wl["vel"].clone() for _ in range(nbuf)]
        gps = [wl["gp"].clone() for _ in range(nbuf)]

        def refill():
            for i in range(nbuf):
                vels[i].copy_(wl["vel"]); gps[i].copy_(wl["gp"])
    else:                       # big slabs: one buffer, regenerated in place between steps
        nbuf = 1
        vels, gps = [wl["vel"]], [wl["gp"]]

        def refill():
            fill_inputs(wl["vel"], wl["gp"], n_glob, wl["zlo"], wl["nz"], ng)
    st = None
    for s_ in range(W):
        with torch.cuda.stream(ctx.stream):
            st = step(vels[s_ % nbuf], gps[s_ % nbuf], wl["p"])
        if nbuf == 1:
            torch.cuda.synchronize(); refill()
    torch.cuda.synchronize()
    refill()
    torch.cuda.synchronize()
    if ctx.nranks > 1:
        dist.barrier()
    t_dev, launches, done, step_ms = 0.0, 0, 0, []
    while done < K:
        chunk = min(nbuf, K - done)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(chunk + 1)]
        torch.cuda.synchronize()
        with torch.cuda.stream(ctx.stream):
            ev[0].record(ctx.stream)
            for i in range(chunk):
                st = step(vels[i], gps[i], wl["p"])
                launches += st.launches
                ev[i + 1].record(ctx.stream)
        torch.cuda.synchronize()
        t_dev += ev[0].elapsed_time(ev[chunk])
        step_ms += [ev[i].elapsed_time(ev[i + 1]) for i in range(chunk)]
        done += chunk
        if done < K:
            refill()   # restore inputs between chunks, outside the event brackets
            torch.cuda.synchronize()
    return allmax(ctx, t_dev) / K, step_ms, st, launches


def eb_block(ctx, with_cpu=True):
    """SURVEY 8(f) rank 4 / BASELINE configs[4], reported next to the headline (not part of `value`): one EB nodal projection
    (b200eb_*) of test_3d/benchmark.channel_cylinder-x scaled to 512 x 128 x 128, device-resident, CUDA-event time of the call;
    `sweep`: the level-0 Gauss-Seidel sweep (8 colour launches) against the measured HBM peak"""
    import numpy as np
    import torch
    from incflo_b200 import eb_geometry as eg, eb_projector as ebp
    n = (512, 128, 128)
    h = 0.4 / n[1]
    geom = eg.cylinder(n, h, 0.05000001, (0.151, 0.2, 0.0), direction=2, small_vfrac=1e-6)
    vel0 = np.zeros((3, n[2] + 2, n[1] + 2, n[0] + 2))
    vel0[0, 1:-1, 1:-1, 1:-1] = (geom.vfrac > 0)
    y = (np.arange(n[1]) + 0.5) / n[1]
    vel0[0, 1:-1, 1:-1, 0] = (6.0 * y * (1.0 - y))[None, :]          # IncfloVelFill, probtype 31
    proj = ebp.EBNodalProjector(n, (h,) * 3, (3, 1, 0), (2, 1, 0), geom.vfrac, geom.intg)
    tv0 = torch.from_numpy(vel0).to(ctx.device)
    phi = torch.zeros((n[2] + 1, n[1] + 1, n[0] + 1), device=ctx.device, dtype=torch.float64)
    times, st = [], None
    for s_ in range(5):
        tv = tv0.clone()
        torch.cuda.synchronize()
        st = proj.project(tv, 1.0, RTOL, ATOL, phi=phi)
        if s_ >= 2:
            times.append(st.ms_total)
    ms = sum(times) / len(times)
    _, nn = proj.level_dims(0)
    nnode = nn[0] * nn[1] * nn[2]
    t_sweep = proj.time_op(0, 0, 1, 10) / 4.0                      # ms per sweep of level 0
    peak = peaks()[0]
    # algorithmic bytes of a sweep: phi of the 4 neighbour colours + rhs + flag in, phi out = 49 B per node with the canonical-row flag
    # (constant sigma, away from the body); 27 coefficients more (265 B) where the row is stored -- here < 2 % of the nodes
    rec = {"n_cell": list(n), "ms_per_projection": ms, "ms_solve": float(st.ms_solve), "vcycles": int(st.iters), "nlevels": int(st.nlevels),
           "resid_over_bnorm": st.resnorm / max(st.rhsnorm, st.resnorm0), "Mcell_updates_per_s": n[0] * n[1] * n[2] / ms / 1e3,
           "launches": int(st.launches), "cut_cells": int(geom.cut_mask().sum()), "covered_cells": int((geom.vfrac == 0).sum()),
           "sweep": {"us": 1e3 * t_sweep, "algorithmic_bytes": 49.0 * nnode, "achieved_GBs": 49.0 * nnode / t_sweep / 1e6,
                     "frac_of_measured_peak": 49.0 * nnode / t_sweep / 1e6 / peak},
           "what": "Hydro::NodalProjector with an EB factory (b200eb_*): channel_cylinder-x, cylinder r = 0.05 along z, mass inflow x-lo (probtype 31), "
                   "pressure outflow x-hi, walls y, periodic z, constant density; initial projection of u = (1, 0, 0)"}
    proj.close()
    del tv0, phi
    torch.cuda.empty_cache()
    if with_cpu:
        # the CPU restatement timed beside it on a bounded sample of the same workload (256 x 64 x 64), and parity at that size
        import time as _t
        from oracle import eb_oracle as eo
        ns = (256, 64, 64)
        hs = 0.4 / ns[1]
        gs = eg.cylinder(ns, hs, 0.05000001, (0.151, 0.2, 0.0), direction=2, small_vfrac=1e-6)
        vs = np.zeros((3, ns[2] + 2, ns[1] + 2, ns[0] + 2))
        vs[0, 1:-1, 1:-1, 1:-1] = (gs.vfrac > 0)
        ys = (np.arange(ns[1]) + 0.5) / ns[1]
        vs[0, 1:-1, 1:-1, 0] = (6.0 * ys * (1.0 - ys))[None, :]
        t0 = _t.perf_counter()
        ref = eo.project(eo.Params(ns, (hs,) * 3, (3, 1, 0), (2, 1, 0)), vs, 1.0, gs.vfrac, gs.intg, RTOL, ATOL)
        tcpu = _t.perf_counter() - t0
        pr2 = ebp.EBNodalProjector(ns, (hs,) * 3, (3, 1, 0), (2, 1, 0), gs.vfrac, gs.intg)
        vg = vs.copy()
        pg = np.zeros((ns[2] + 1, ns[1] + 1, ns[0] + 1))
        st2 = pr2.project(vg, 1.0, RTOL, ATOL, phi=pg)
        pr2.close()
        rl2 = lambda a, b: float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))
        rec["cpu_port"] = {"n_cell": list(ns), "seconds": tcpu, "Mcell_updates_per_s": ns[0] * ns[1] * ns[2] / tcpu / 1e6, "cores": 1,
                           "vcycles": int(ref["info"]["iters"]), "what": "oracle/eb_oracle.py (numpy + SciPy sparse, one thread): CPU restatement, not AMReX"}
        rec["parity"] = {"n_cell": list(ns), "rel_l2_phi": rl2(pg[:ns[2]], ref["phi"]), "rel_l2_u": rl2(vg[:, 1:-1, 1:-1, 1:-1], ref["vel"]),
                         "vcycles_gpu": int(st2.iters), "vcycles_oracle": int(ref["info"]["iters"]), "tolerance": 1e-9}
        if not (rec["parity"]["rel_l2_phi"] < 1e-9 and rec["parity"]["rel_l2_u"] < 1e-9):
            rec["parity"]["FAILED"] = True
    return rec


def mac_block(ctx, n1):
    """SURVEY 8(f) rank 3, reported next to the headline (not part of `value`): one MAC projection (Hydro::MacProjector
    semantics, b200mac_*) of an n1^3 rayleigh_taylor-like field, device-resident, CUDA-event time of the call"""
    import torch
    from incflo_b200 import mac_projector as mp
    dev = ctx.device
    g = torch.Generator(device=dev); g.manual_seed(7)
    z = (torch.arange(n1, device=dev, dtype=torch.float64) + 0.5) / n1
    rho = (1.0 + 1.5 * (1.0 + torch.tanh((z - 0.5) / 0.05)))[:, None, None].expand(n1, n1, n1).contiguous()
    dt = 0.01
    bx = dt / (0.5 * (rho + torch.roll(rho, 1, 2))); bx = torch.cat([bx, bx[:, :, :1]], 2).contiguous()
    by = dt / (0.5 * (rho + torch.roll(rho, 1, 1))); by = torch.cat([by, by[:, :1]], 1).contiguous()
    rz = torch.cat([rho[:1], rho, rho[-1:]], 0)
    bz = (dt / (0.5 * (rz[:-1] + rz[1:]))).contiguous()

    def smooth(a):
        for ax in range(3):
            a = 0.5 * a + 0.25 * (torch.roll(a, 1, ax) + torch.roll(a, -1, ax))
        return a
    u0 = smooth(torch.randn((n1, n1, n1 + 1), device=dev, dtype=torch.float64, generator=g)); u0[:, :, -1] = u0[:, :, 0]
    v0 = smooth(torch.randn((n1, n1 + 1, n1), device=dev, dtype=torch.float64, generator=g)); v0[:, -1] = v0[:, 0]
    w0 = smooth(torch.randn((n1 + 1, n1, n1), device=dev, dtype=torch.float64, generator=g)); w0[0] = 0; w0[-1] = 0
    proj = mp.MacProjector((n1, n1, n1), (1.0 / n1,) * 3, (0, 0, 1), (0, 0, 1))
    proj.updateCoeffs([bx, by, bz])
    phi = torch.zeros((n1, n1, n1), device=dev, dtype=torch.float64)
    times, st = [], None
    for s_ in range(5):
        u, v, w = u0.clone(), v0.clone(), w0.clone()
        torch.cuda.synchronize()
        st = proj.project(u, v, w, RTOL, ATOL, mac_phi=phi)
        if s_ >= 2:
            times.append(st.ms_total)
    div = (u[:, :, 1:] - u[:, :, :-1]) * n1 + (v[:, 1:] - v[:, :-1]) * n1 + (w[1:] - w[:-1]) * n1
    rec = {"n_cell": [n1] * 3, "ms_per_projection": sum(times) / len(times), "vcycles": int(st.iters),
           "resid_over_bnorm": st.resnorm / max(st.rhsnorm, st.resnorm0), "max_div_over_bnorm": float((div - div.mean()).abs().max()) / max(st.rhsnorm, st.resnorm0),
           "Mcell_updates_per_s": n1 ** 3 / (sum(times) / len(times)) / 1e3, "launches": int(st.launches),
           "what": "Hydro::MacProjector / MLABecLaplacian semantics (b200mac_*), variable beta = dt/rho on faces, periodic x/y + walls z"}
    proj.close()
    del u0, v0, w0, bx, by, bz, rho, phi
    torch.cuda.empty_cache()
    return rec


def north_star_block(ctx):
    """BASELINE north_star's single-GPU target, measured in the default run so that it is in the driver's record: the 512^3
    variable-density projection to rtol 1e-11 on one B200, with the level-0 smoother / residual kernels against the HBM peak"""
    import torch
    from incflo_b200 import nodal_projector as npj
    n = (512, 512, 512)
    wl = workload(n, 1, 0, ctx.device)
    proj = make_projection(ctx, n, nranks=1)
    # the two kernels first, timed alone after one warm-up solve (MEASURED_PEAKS' hbm_gbs is a burst figure as well); then the solves,
    # which run into the power cap at this size (clocks sampled over both)
    sampler = ClockSampler(ctx.local)
    sampler.start()
    timed_solves(ctx, proj, wl, 1, 0, n, "clone")
    _, nn = proj.level_dims(0)
    nodes = nn[0] * nn[1] * nn[2]
    ms_sm = proj.time_op(0, npj.OP_SMOOTH, 2, reps=20) / 2.0
    ms_res = proj.time_op(0, npj.OP_RESIDUAL, 0, reps=20)
    ms, _, st, _ = timed_solves(ctx, proj, wl, 3, 1, n, "clone")
    clocks = sampler.stop()
    peak, _ = peaks()
    gbs = lambda t: 32.0 * nodes / (t * 1e-3) / 1e9
    solve_bytes = solve_roofline(nodes, n[0] * n[1] * n[2], st.iters, True)
    rec = {"n_cell": list(n), "ms_per_solve": ms, "vcycles": int(st.iters), "resid_over_bnorm": st.resnorm / max(st.rhsnorm, st.resnorm0),
           "Mcell_updates_per_s": n[0] * n[1] * n[2] / ms / 1e3, "steps": 3, "warmup": 1,
           "smoother_sweep": {"us": ms_sm * 1e3, "GBs": gbs(ms_sm), "frac_of_measured_peak": gbs(ms_sm) / peak, "frac_of_nominal_8TBs": gbs(ms_sm) / 8000.0},
           "residual": {"us": ms_res * 1e3, "GBs": gbs(ms_res), "frac_of_measured_peak": gbs(ms_res) / peak, "frac_of_nominal_8TBs": gbs(ms_res) / 8000.0},
           "whole_solve_frac_of_measured_peak": solve_bytes / (ms * 1e-3) / 1e9 / peak, "clocks": clocks,
           "target": "north_star: 512^3 variable density to rtol 1e-11 on one B200, smoother / residual kernels >= 60 % of HBM peak"}
    proj.close()
    del wl
    torch.cuda.empty_cache()
    return rec


def strong_block(ctx, sizes):
    """strong scaling: the n^3 problem on all N GPUs vs the SAME solve measured on one GPU (rank 0) in this run"""
    import torch
    import torch.distributed as dist
    out = {}
    for n1 in sizes:
        n = (n1, n1, n1)
        rec = {"n_cell": list(n)}
        restore = "clone" if n1 <= 512 else "regen"
        K, W = (3, 2) if n1 <= 512 else (2, 1)
        # one GPU (rank 0 alone; the other ranks wait at the barrier)
        if ctx.rank == 0:
            try:
                wl = workload(n, 1, 0, ctx.device)
                solo = Ctx(); solo.__dict__.update(ctx.__dict__); solo.nranks = 1
                proj = make_projection(solo, n, nranks=1)
                ms1, _, st1, _ = timed_solves(solo, proj, wl, K, W, n, restore)
                proj.close()
                rec.update(ms_per_solve_1gpu=ms1, vcycles_1gpu=int(st1.iters))
                del wl, proj
            except Exception as e:   # e.g. out of memory for the one-GPU 1024^3 problem next to other tenants
                rec.update(ms_per_solve_1gpu=None, error_1gpu=repr(e)[:200])
            torch.cuda.empty_cache()
        dist.barrier()
        wl = workload(n, ctx.nranks, ctx.rank, ctx.device)
        proj = make_projection(ctx, n)
        msN, _, stN, _ = timed_solves(ctx, proj, wl, K, W, n, restore)
        rec.update(ms_per_solve=msN, vcycles=int(stN.iters), resid_over_bnorm=stN.resnorm / max(stN.rhsnorm, stN.resnorm0),
                   Mcell_updates_per_s=n1 ** 3 / msN / 1e3, steps=K, warmup=W)
        proj.close()
        del wl, proj
        torch.cuda.empty_cache()
        if ctx.rank == 0 and rec.get("ms_per_solve_1gpu"):
            rec["speedup_vs_1gpu"] = rec["ms_per_solve_1gpu"] / msN
        out[str(n1)] = rec
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from incflo_b200 import nodal_projector as npj

    ctx = Ctx()
    ctx.nranks = nranks = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.rank = rank = int(os.environ.get("RANK", "0"))
    ctx.local = local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the product path has no CPU fallback"
    torch.cuda.set_device(local)
    ctx.device = device = f"cuda:{local}"
    if nranks > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    ctx.stream = torch.cuda.Stream()   # the launching stream: the handle runs on it, the events are recorded on it
    N, K, W = args.n, args.steps, args.warmup
    strong = args.scaling == "strong" and nranks > 1
    threads = oracle_threads() if rank == 0 else 0

    # ---------------- parity first: a fast wrong answer is not a result ----------------
    parity = None
    if not args.no_parity:
        parity = parity_block(ctx)
        if not parity["ok"]:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "error": "parity check failed", "parity": parity}))
            sys.stdout.flush()
            sys.exit(3)

    # ---------------- device-resident arm ----------------
    n_glob = (N, N, N) if (strong or nranks == 1) else (N, N, N * nranks)   # weak (default): an N^3 slab per rank
    wl = workload(n_glob, nranks, rank, device)
    proj = make_projection(ctx, n_glob)
    transport = proj.halo_transport()
    pmap = proj.peer_map()
    ncell = n_glob[0] * n_glob[1] * n_glob[2] // nranks   # cells per rank
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    ms_per_step, step_ms, st, launches = timed_solves(ctx, proj, wl, K, W, n_glob, "clone" if N <= 512 else "regen")
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    iters = st.iters
    value = ncell * nranks / (ms_per_step * 1e-3) / 1e6
    resid_ratio = st.resnorm / max(st.rhsnorm, st.resnorm0)

    # ---------------- roofline of the dominant kernel (level-0 smoother sweep) ----------------
    _, nn = proj.level_dims(0)
    nodes = nn[0] * nn[1] * nn[2]
    reps = 20
    ms_sm = proj.time_op(0, npj.OP_SMOOTH, 2, reps=reps) / 2.0   # CUDA events around back-to-back launches
    ms_res = proj.time_op(0, npj.OP_RESIDUAL, 0, reps=reps)
    peak, peak_src = peaks()
    alg_bytes = 32.0 * nodes
    achieved = alg_bytes / (ms_sm * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f).get(f"k_smooth_iso_var_{N}")
    except Exception:
        pass
    sweeps_per_step = iters * 16
    roofline = {"bound": "hbm", "kernel": "level-0 Gauss-Seidel sweep, variable sigma (one launch = one sweep over the level)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": "profiles/ncu_traffic.json (committed ncu --set full capture of this kernel at this size)",
                "frac_of_nominal_8TBs": achieved / 8000.0,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "us_per_launch": ms_sm * 1e3,
                "launches_per_step": sweeps_per_step, "share_of_step": sweeps_per_step * ms_sm / ms_per_step,
                "residual_kernel": {"us_per_launch": ms_res * 1e3, "achieved": alg_bytes / (ms_res * 1e-3) / 1e9,
                                    "frac": alg_bytes / (ms_res * 1e-3) / 1e9 / peak}}
    # the whole projection against the same peak (per rank: local nodes / cells, the realised V-cycle count)
    solve_bytes = solve_roofline(nodes, ncell, iters, True)
    roofline["whole_solve"] = {"algorithmic_bytes": solve_bytes, "achieved": solve_bytes / (ms_per_step * 1e-3) / 1e9,
                               "frac": solve_bytes / (ms_per_step * 1e-3) / 1e9 / peak}

    # ---------------- e2e arm: host (pinned) buffers through the same C-ABI call ----------------
    ng = wl["ng"]

    def step(vel, gp, p, rho):
        return proj.apply_nodal_projection(vel, ng, gp, p, density=rho, ngd=ng, scaling_factor=wl["dt"],
                                           mg_rtol=RTOL, mg_atol=ATOL)
    Ke = 0 if args.no_e2e else min(K, 3)
    hv = [wl["vel"].cpu().pin_memory() for _ in range(Ke)]
    hg = [wl["gp"].cpu().pin_memory() for _ in range(Ke)]
    hp = [wl["p"].cpu().pin_memory() for _ in range(Ke)]
    hr = wl["rho"].cpu().pin_memory()
    if Ke:
        step(hv[0].numpy(), hg[0].numpy(), hp[0].numpy(), hr.numpy())          # warm-up (allocates staging buffers)
        hv[0].copy_(wl["vel"].cpu()); hg[0].copy_(wl["gp"].cpu())
    torch.cuda.synchronize()
    if nranks > 1:
        dist.barrier()
    te = 0.0
    h2d = d2h = 0
    for i in range(Ke):
        a = time.perf_counter()
        ste = step(hv[i].numpy(), hg[i].numpy(), hp[i].numpy(), hr.numpy())
        te += time.perf_counter() - a
        h2d, d2h = ste.h2d_bytes, ste.d2h_bytes
    te = allmax(ctx, te)
    e2e_val = ncell * nranks / (te / Ke) / 1e6 if Ke else None
    proj.close()
    del wl, hv, hg, hp, hr
    torch.cuda.empty_cache()

    # ---------------- MAC projection record (N = 1): the next operator of SURVEY 8(f), same size ----------------
    mac_rec = None
    if nranks == 1 and not args.no_mac and N <= 256:
        try:
            mac_rec = mac_block(ctx, N)
        except Exception as e:   # never let the extra record take the headline down
            mac_rec = {"error": repr(e)[:200]}

    # ---------------- north_star's single-GPU target (N = 1): 512^3 ----------------
    ns_rec = None
    if nranks == 1 and not args.no_512 and N == 256:
        try:
            ns_rec = north_star_block(ctx)
        except Exception as e:
            ns_rec = {"error": repr(e)[:200]}

    # ---------------- EB nodal projection record (N = 1): BASELINE configs[4] ----------------
    eb_rec = None
    if nranks == 1 and not args.no_eb:
        try:
            eb_rec = eb_block(ctx, with_cpu=not args.no_cpu)
        except Exception as e:
            eb_rec = {"error": repr(e)[:200]}

    # ---------------- strong-scaling record (N > 1): 512^3 / 1024^3 in total vs one GPU, measured here ----------------
    strong_rec = None
    if nranks > 1 and not args.no_strong and not strong:
        sizes = [s for s in (512, 1024) if s % nranks == 0]
        strong_rec = strong_block(ctx, sizes)

    if rank == 0:
        tname = {0: "none (1 GPU)", 1: "peer memory (" + {1: "cuMem + POSIX fd", 2: "cudaIpc"}.get(pmap, "?") + ")", 2: "nccl"}[transport]
        par = "1 GPU" if nranks == 1 else (
            f"z-slab decomposition over {nranks} GPUs; halo planes: "
            + ("stores / loads into the neighbours' arenas over NVLink peer memory, issued by the solver kernels"
               if transport == 1 else "FALLBACK grouped ncclSend/ncclRecv (peer mapping unavailable)")
            + f"; ncclAllReduce for norms / solvability; domain {n_glob[0]}x{n_glob[1]}x{n_glob[2]}")
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": nranks, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"rayleigh_taylor variable-density (sigma=dt/rho, 4:1) {N}^3 {'in total' if strong else 'per GPU'}, periodic x/y + "
                                       f"walls z, nodal projection to rtol 1e-11 (BASELINE configs[1])",
                           "rtol": RTOL, "atol": ATOL, "vcycles": iters, "resid_over_bnorm": resid_ratio,
                           "cycle": "V(2,2) x 4 sweeps (reference defaults)", "inputs_vs_l2": "working set >> 126 MB L2",
                           "parallelism": par, "halo_transport": tname,
                           "solves_per_s": 1e3 / ms_per_step,   # projections of the whole (global) domain per second
                           "ms_per_step_median": sorted(step_ms)[len(step_ms) // 2], "ms_per_step_min": min(step_ms)},
                "clocks": clocks, "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                                          "d2h_bytes_per_step": int(d2h), "steps": Ke, "ms_per_step": te / Ke * 1e3 if Ke else None},
                "gpu_launches": int(launches), "roofline": roofline, "parity": parity, "wall_s_timed_region": wall}
        if strong_rec is not None:
            line["strong"] = strong_rec
        if mac_rec is not None:
            line["mac_projection"] = mac_rec
        if eb_rec is not None:
            line["eb_projection"] = eb_rec
        if ns_rec is not None:
            line["north_star_512"] = ns_rec
        if nranks == 1 and not args.no_cpu:
            times, it_cpu, _ = cpu_port_run((N, N, N), 1, 0, "reference")
            tcpu = min(times)
            line["cpu_baseline"] = {"value": N ** 3 / tcpu / 1e6, "unit": UNIT, "cores": threads, "kind": "port", "host_cpus": os.cpu_count(),
                                    "sample": f"one full {N}^3 projection of the same workload ({tcpu:.1f} s, {it_cpu} V-cycles) with the "
                                              f"reference's CPU algorithm (AMReX multi-box semantics: lexicographic Gauss-Seidel in "
                                              f"{min(REF_BOX, N // 2)}^3 boxes, 4 sweeps per smooth call without halo refresh); CPU restatement, "
                                              f"not incflo/AMReX"}
        print(json.dumps(line))
    if nranks > 1:
        dist.destroy_process_group()


def _json_only_stdout():
    """Native libraries (NCCL's version banner) write to fd 1; the contract is ONE JSON line on
    stdout.  Point fd 1 at stderr for the run and keep the real stdout for the final line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--size", dest="n", type=int, default=256,
                    help="cells per side (--size: spelling that torchrun's own option parser does not claim)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = n^3 cells per GPU (default, the driver's scaling run); strong = n^3 in total")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="tuning sweeps only: skip the host-buffer arm (the line is then not a valid bench line)")
    ap.add_argument("--no-parity", action="store_true", help="tuning sweeps only: skip the oracle check before timing")
    ap.add_argument("--no-mac", action="store_true", help="N = 1: skip the MAC projection record")
    ap.add_argument("--no-512", dest="no_512", action="store_true", help="N = 1: skip the 512^3 north_star record")
    ap.add_argument("--no-eb", action="store_true", help="N = 1: skip the EB nodal projection record (BASELINE configs[4])")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the 512^3 / 1024^3 strong-scaling record")
    args = ap.parse_args()
    real_stdout = _json_only_stdout()
    sys.stdout = real_stdout
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    real_stdout.flush()


if __name__ == "__main__":
    main()
