"""Per-level smoother / residual timings of the slab-decomposed path (collective: run under torchrun).
Usage: torchrun ... tools/kernel_bench_dist.py [N]   -- N^3 cells per rank (weak), rayleigh_taylor"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from incflo_b200 import nodal_projector as npj

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
wl = bench.workload((N, N, N * world), world, rank, dev)
idt = torch.zeros(128, dtype=torch.uint8, device=dev)
if rank == 0:
    idt.copy_(torch.tensor(list(npj.nccl_unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
proj = npj.IncfloProjection(wl["n"], wl["dx"], wl["bclo"], wl["bchi"], device=local, rank=rank, nranks=world,
                            nccl_id=bytes(idt.cpu().tolist()))
dbg = int(os.environ.get("B200NP_DBG_HALO", "0"))
st = proj.stats if dbg else proj.apply_nodal_projection(wl["vel"].clone(), wl["ng"], wl["gp"].clone(), wl["p"].clone(), density=wl["rho"], ngd=wl["ng"],
                                 scaling_factor=wl["dt"])
if dbg:   # no solve (halos are deliberately wrong): sigma still has to be in place for the variable-sigma kernels
    import ctypes as C
    ng, Nn = wl["ng"], N
    sig = (wl["dt"] / wl["rho"][ng:ng + Nn, ng:ng + Nn, ng:ng + Nn]).contiguous()
    ptr, box, _ = npj._ptr_box(sig, (0, 0, proj.zlo), 1)
    assert proj._L.b200np_set_sigma(proj._h, ptr, C.byref(box), 1.0) == 0
nl = proj._L.b200np_nlevels(proj._h)
out = [f"rank {rank}: iters={st.iters} ms_total={st.ms_total:.2f} env DBG_HALO={os.environ.get('B200NP_DBG_HALO', '0')} "
       f"FUSE={os.environ.get('B200NP_FUSE_HALO', '1')} P2P={os.environ.get('B200NP_P2P', '1')}"]
for lev in range(min(nl, 5)):
    n, nn = proj.level_dims(lev)
    ms8 = proj.time_op(lev, npj.OP_SMOOTH, 8, reps=10)
    msr = proj.time_op(lev, npj.OP_RESIDUAL, 0, reps=10)
    out.append(f"  lev{lev} local nodes {nn}: 8 sweeps {ms8 * 1e3:8.1f} us ({ms8 * 1e3 / 8:6.1f} us/sweep)  residual {msr * 1e3:7.1f} us")
if rank == 0:
    print("\n".join(out), flush=True)
proj.close()
dist.destroy_process_group()
