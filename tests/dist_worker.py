"""Worker of the multi-GPU parity test (launched by torchrun, one process per GPU):
slab-decomposed projection vs the CPU oracle on the global problem."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from incflo_b200 import nodal_projector as npj, problems, slab
    from oracle import pyoracle as po
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))

    def fresh_id():   # an ncclUniqueId serves exactly one communicator
        idt = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local}")
        if rank == 0:
            idt.copy_(torch.tensor(list(npj.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        return bytes(idt.cpu().tolist())

    ok = True
    for cfgname, N, ng in (("tgv", 64, 2), ("rt", 64, 3)):
        n = (N, N, N)
        dx = (1.0 / N,) * 3
        bclo = bchi = problems.CONFIGS[cfgname]["bclo"]
        var = problems.CONFIGS[cfgname]["var"]
        rng = np.random.default_rng(11)
        if cfgname == "tgv":
            vel = problems.taylor_green(n, ng, "cpu").numpy().copy()
        else:
            vel = problems.rayleigh_taylor_velocity(n, ng, "cpu", "b").numpy().copy()
        rho = problems.rayleigh_taylor_density(n, ng, "cpu").numpy().copy() if var else None
        gp = 0.1 * rng.standard_normal((3, N, N, N))
        p = np.zeros((N + 1, N + 1, N + 1))
        dt = 0.45 / N
        # oracle on the global problem (same tile ordering; chunks are slab-local on the GPUs, so the
        # comparison is to tolerance, not sweep-by-sweep)
        ov, ogp, op_ = vel.copy(), gp.copy(), p.copy()
        prm = po.make_params(n, dx, bclo, bchi, smoother=po.SM_BOX, box=(64, 16, 64), box_order=po.SM_PLANE4,
                             box_stale_per_call=0)
        status, ost = po.apply_nodal_projection(prm, ov, ng, ogp, op_, density=rho, ngd=ng, ro_0=1.0, scaling_factor=dt)
        assert status == 0
        # my slab
        clo, chi, nlo, nhi = slab.slab_range(n, bclo, rank, world)
        lv = torch.from_numpy(slab.cut(vel, clo, chi, ng)).cuda()
        lr = torch.from_numpy(slab.cut(rho, clo, chi, ng)).cuda() if var else None
        lg = torch.from_numpy(gp[:, clo:chi + 1].copy()).cuda()
        lp = torch.from_numpy(p[clo:chi + 2].copy()).cuda()
        ip = npj.IncfloProjection(n, dx, bclo, bchi, device=local, rank=rank, nranks=world, nccl_id=fresh_id())
        st = ip.apply_nodal_projection(lv, ng, lg, lp, density=lr, ngd=ng, ro_0=1.0, scaling_factor=dt)
        torch.cuda.synchronize()
        gv, gg, gpn = lv.cpu().numpy(), lg.cpu().numpy(), lp.cpu().numpy()
        inner = (slice(None), slice(ng, -ng), slice(ng, -ng), slice(ng, -ng))
        rv = slab.cut(ov, clo, chi, ng)[inner]
        ev = np.linalg.norm(gv[inner] - rv) / max(np.linalg.norm(rv), 1e-300)
        eg = np.linalg.norm(gg - ogp[:, clo:chi + 1]) / np.linalg.norm(ogp[:, clo:chi + 1])
        # pressure: compare after removing the GLOBAL mean -> gather sums
        dp_ = gpn[:-1] - op_[clo:chi + 1]     # planes owned by this slab
        t = torch.tensor([dp_.sum(), dp_.size, 0.0, 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        mean = (t[0] / t[1]).item()
        t2 = torch.tensor([((dp_ - mean) ** 2).sum(), (op_[clo:chi + 1] ** 2).sum()], dtype=torch.float64, device="cuda")
        dist.all_reduce(t2)
        ep = float(torch.sqrt(t2[0] / t2[1]).item())
        good = (st.status == 0 and abs(st.iters - ost.iters) <= 1 and ev < 1e-9 and eg < 1e-9 and ep < 1e-9)
        print(f"[rank {rank}] {cfgname}: iters gpu={st.iters} oracle={ost.iters} rel-L2 vel={ev:.2e} gp={eg:.2e} p={ep:.2e} "
              f"{'ok' if good else 'FAIL'}", flush=True)
        ok = ok and good
        ip.close()
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST OK" if flag.item() == 1.0 else "DIST FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
