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
roofline: the dominant kernel (tile-resident Gauss-Seidel sweep k_smooth_iso, level 0): algorithmic
         bytes (32 B/node variable sigma) / CUDA-event time per launch vs MEASURED_PEAKS.json hbm_gbs
         (burst copy figure; the kernel is timed alone, back to back); traffic = dram bytes of one
         launch from the committed ncu --set full capture (profiles/ncu_traffic.json).
cpu_baseline: the CPU oracle (a port of the reference algorithm, NOT incflo/AMReX itself, which
         cannot be built offline) on the box's host cores, on a bounded sample (128^3 of the same
         workload).
--impl reference: times that CPU port alone (the reference's own CPU build needs the un-vendored
         AMReX + AMReX-Hydro and MPI, none of which exist offline).
N > 1  : z-slab decomposition, one process per GPU (torchrun), weak scaling: every rank owns
         256 x 256 x 256 cells of a 256 x 256 x 256N domain.
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


def workload(N, nranks, rank, device, ng=3, strong=False):
    """per-rank slab of the rayleigh_taylor workload.  weak scaling: N x N x N cells per rank
    (domain N x N x N*nranks); strong scaling: the N^3 domain cut into nranks z slabs."""
    import torch
    from incflo_b200 import problems
    n_glob = (N, N, N) if strong else (N, N, N * nranks)
    nz = N // nranks if strong else N
    dt = 0.45 / N
    with problems.slab(rank * nz, nz):   # this rank's cell planes [rank*nz, (rank+1)*nz) of the global closed forms
        vel = problems.rayleigh_taylor_velocity(n_glob, ng, device, "b")
        rho = problems.rayleigh_taylor_density(n_glob, ng, device)
    gp = torch.zeros((3, nz, N, N), dtype=torch.float64, device=device)
    gp[2] = -0.05  # a hydrostatic-like old pressure gradient so the pre-add does work
    p = torch.zeros((nz + 1, N + 1, N + 1), dtype=torch.float64, device=device)
    return dict(n=n_glob, dx=(1.0 / N,) * 3, dt=dt, vel=vel, rho=rho, gp=gp, p=p, ng=ng,
                bclo=(0, 0, 1), bchi=(0, 0, 1))


def cpu_port_run(N, steps, warmup):
    """the CPU oracle (port of the reference algorithm) on a bounded sample; returns list of seconds"""
    import numpy as np
    from incflo_b200 import problems
    from oracle import pyoracle as po
    ng = 3
    n = (N, N, N)
    vel0 = problems.rayleigh_taylor_velocity(n, ng, "cpu", "b").numpy()
    rho = problems.rayleigh_taylor_density(n, ng, "cpu").numpy()
    gp0 = np.zeros((3, N, N, N)); gp0[2] = -0.05
    prm = po.make_params(n, (1.0 / N,) * 3, (0, 0, 1), (0, 0, 1), smoother=po.SM_BOX, box=(64, 16, 64),
                         box_order=po.SM_PLANE4, box_stale_per_call=0)
    times, iters = [], 0
    for s in range(warmup + steps):
        vel = vel0.copy(); gp = gp0.copy(); p = np.zeros((N + 1,) * 3)
        t0 = time.perf_counter()
        status, st = po.apply_nodal_projection(prm, vel, ng, gp, p, density=rho, ngd=ng, scaling_factor=0.45 / N,
                                               rtol=RTOL, atol=ATOL)
        dt = time.perf_counter() - t0
        assert status == 0
        iters = st.iters
        if s >= warmup:
            times.append(dt)
    return times, iters


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    Ns = 128  # bounded sample: 1/8 of the 256^3 workload per step
    cores = os.cpu_count()
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    steps, warmup = min(args.steps, 5), min(args.warmup, 1)
    times, iters = cpu_port_run(Ns, steps, warmup)
    t = sum(times) / len(times)
    val = Ns ** 3 / t / 1e6
    sample = f"{Ns}^3 rayleigh_taylor variable-density projection per step (1/8 of the {args.n}^3 workload), {len(times)} steps"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
            "warmup": warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"rayleigh_taylor variable-density {args.n}^3 nodal projection (BASELINE configs[1]); "
                                   f"CPU arm runs the bounded {Ns}^3 sample", "rtol": RTOL, "atol": ATOL, "vcycles": iters},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "note": "CPU restatement of AMReX MLMG nodal projection (OpenMP); incflo/AMReX itself cannot be "
                                     "built offline (AMReX, AMReX-Hydro, MPI not vendored)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from incflo_b200 import nodal_projector as npj

    nranks = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the product path has no CPU fallback"
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if nranks > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    N, K, W = args.n, args.steps, args.warmup
    strong = args.scaling == "strong" and nranks > 1
    wl = workload(N, nranks, rank, device, strong=strong)   # weak (default): every rank owns an N^3 slab of an N x N x (N*nranks) domain
    nccl_id = None
    if nranks > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device=device)
        if rank == 0:
            idt.copy_(torch.tensor(list(npj.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().tolist())
    ng = wl["ng"]
    proj = npj.IncfloProjection(wl["n"], wl["dx"], wl["bclo"], wl["bchi"], device=local, rank=rank, nranks=nranks,
                                nccl_id=nccl_id)
    stream = torch.cuda.Stream()           # the launching stream: the handle runs on it, the events are recorded on it
    proj.set_stream(stream.cuda_stream)
    ncell = N ** 3 // nranks if strong else N ** 3   # cells per rank

    def step(vel, gp, p, rho):
        return proj.apply_nodal_projection(vel, ng, gp, p, density=rho, ngd=ng, scaling_factor=wl["dt"],
                                           mg_rtol=RTOL, mg_atol=ATOL)

    # ---------------- device-resident arm ----------------
    nbuf = min(K, 16)                      # every timed step gets its own untouched inputs
    vels = [wl["vel"].clone() for _ in range(nbuf)]
    gps = [wl["gp"].clone() for _ in range(nbuf)]
    ps = [wl["p"].clone() for _ in range(nbuf)]

    def refill():
        for i in range(nbuf):
            vels[i].copy_(wl["vel"]); gps[i].copy_(wl["gp"])
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for s_ in range(W):
            st = step(vels[s_ % nbuf], gps[s_ % nbuf], ps[s_ % nbuf], wl["rho"])
    torch.cuda.synchronize()
    refill()
    if nranks > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_dev = 0.0
    launches = 0
    iters = 0
    t0 = time.perf_counter()
    done = 0
    step_ms = []
    while done < K:
        chunk = min(nbuf, K - done)
        # one event bracket around the chunk (the reported time) + one event after every step (median / min)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(chunk + 1)]
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            ev[0].record(stream)
            for i in range(chunk):
                st = step(vels[i], gps[i], ps[i], wl["rho"])
                launches += st.launches
                iters = st.iters
                ev[i + 1].record(stream)
        torch.cuda.synchronize()
        t_dev += ev[0].elapsed_time(ev[chunk])
        step_ms += [ev[i].elapsed_time(ev[i + 1]) for i in range(chunk)]
        done += chunk
        if done < K:
            refill()   # restore inputs between chunks, outside the event brackets
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    if nranks > 1:
        tt = torch.tensor([t_dev], dtype=torch.float64, device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev = float(tt.item())
    ms_per_step = t_dev / K
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
    roofline = {"bound": "hbm", "kernel": "k_smooth_iso<variable sigma> level 0 (one Gauss-Seidel sweep)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "us_per_launch": ms_sm * 1e3,
                "launches_per_step": sweeps_per_step, "share_of_step": sweeps_per_step * ms_sm / ms_per_step,
                "residual_kernel": {"us_per_launch": ms_res * 1e3, "achieved": alg_bytes / (ms_res * 1e-3) / 1e9,
                                    "frac": alg_bytes / (ms_res * 1e-3) / 1e9 / peak}}
    # the whole projection against the same peak (per rank: local nodes / cells, the realised V-cycle count)
    solve_bytes = solve_roofline(nodes, ncell, iters, True)
    roofline["whole_solve"] = {"algorithmic_bytes": solve_bytes, "achieved": solve_bytes / (ms_per_step * 1e-3) / 1e9,
                               "frac": solve_bytes / (ms_per_step * 1e-3) / 1e9 / peak}

    # ---------------- e2e arm: host (pinned) buffers through the same C-ABI call ----------------
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
    if nranks > 1:
        tt = torch.tensor([te], dtype=torch.float64, device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        te = float(tt.item())
    e2e_val = ncell * nranks / (te / Ke) / 1e6 if Ke else None

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": nranks, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"rayleigh_taylor variable-density (sigma=dt/rho, 4:1) {N}^3 {'in total' if strong else 'per GPU'}, periodic x/y + "
                                       f"walls z, nodal projection to rtol 1e-11 (BASELINE configs[1])",
                           "rtol": RTOL, "atol": ATOL, "vcycles": iters, "resid_over_bnorm": resid_ratio,
                           "cycle": "V(2,2) x 4 sweeps (reference defaults)", "inputs_vs_l2": "working set >> 126 MB L2",
                           "parallelism": "1 GPU" if nranks == 1 else f"z-slab decomposition over {nranks} GPUs (NCCL halo planes + allreduce), "
                                                                                 f"domain {N}x{N}x{N if strong else N * nranks}",
                           "solves_per_s": 1e3 / ms_per_step,   # projections of the whole (global) domain per second
                           "ms_per_step_median": sorted(step_ms)[len(step_ms) // 2], "ms_per_step_min": min(step_ms)},
                "clocks": clocks, "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                                          "d2h_bytes_per_step": int(d2h), "steps": Ke, "ms_per_step": te / Ke * 1e3 if Ke else None},
                "gpu_launches": int(launches), "roofline": roofline, "wall_s_timed_region": wall}
        if nranks == 1 and not args.no_cpu:
            cores = os.cpu_count()
            os.environ.setdefault("OMP_NUM_THREADS", str(cores))
            Ns = 128
            times, it_cpu = cpu_port_run(Ns, 2, 0)
            tcpu = min(times)
            line["cpu_baseline"] = {"value": Ns ** 3 / tcpu / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{Ns}^3 of the same workload (1/8 of the cells), best of {len(times)} runs, "
                                              f"{it_cpu} V-cycles; CPU restatement of the AMReX algorithm, not incflo/AMReX"}
        print(json.dumps(line))
    proj.close()
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
