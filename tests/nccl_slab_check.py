"""Multi-GPU check of the slab path with real NCCL (launch with torchrun, one rank per GPU):
condense on every rank -> cut-plane exchange (NCCL send/recv behind the C ABI) -> owner-computes assembly of the owned columns,
gathered on rank 0 and compared bit-for-bit with the single-GPU global assembly; then the lambda all-gather and
the backward step against the oracle.  Used by tests/test_gpu_multi.py and by hand:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/nccl_slab_check.py 6 5 8
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gridaphybrid_b200 as gh  # noqa: E402
from gridaphybrid_b200.distributed import SlabAssembler, SlabLayout  # noqa: E402


def main():
    gdims = tuple(int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (6, 5, 8)
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = gh.Context(local)
    ctx.comm_init_from_torch()                       # the library's own NCCL communicator (ghb_comm_init); torch ships the id
    ndofs, touched = [30, 4, 36], np.ones((3, 3), bool)
    plan = ctx.plan_blocks(ndofs, touched, [1, 2], [3])
    L = SlabLayout(gdims, 6, rank, world)
    dev = torch.device("cuda", local)
    n = L.ncells
    A = torch.empty((n, plan.lenA), dtype=torch.float64, device=dev)
    b = torch.empty((n, plan.lenb), dtype=torch.float64, device=dev)
    ctx.synth_fill(plan, L.cell_start, n, A, b)
    S = torch.empty((n, 36 * 36), dtype=torch.float64, device=dev)
    g = torch.empty((n, 36), dtype=torch.float64, device=dev)
    info = torch.empty(n, dtype=torch.int32, device=dev)
    ctx.condense(plan, n, A, b, S, g, info)
    L0 = SlabLayout(gdims, 6, 0, 1)
    ids_all = L0.cell_dof_ids(torch.arange(L0.ncells_global, device=dev))
    ndir = int((-ids_all).max().item())
    dv = torch.linspace(-1.0, 1.0, ndir, dtype=torch.float64, device=dev)
    asm = SlabAssembler(ctx, gdims, 6, rank, world, dirichlet_values=dv)
    nz = torch.empty(asm.nnz, dtype=torch.float64, device=dev)
    rhs = torch.empty(asm.nrows_local, dtype=torch.float64, device=dev)
    asm.assemble(S, g, nz, rhs)                      # cut-plane exchange through the C ABI (ghb_exchange_cut_plane_f64)
    colptr, rowval = asm.pattern()
    torch.cuda.synchronize()
    parts = [None] * world
    dist.all_gather_object(parts, dict(colptr=colptr.cpu().numpy(), rowval=rowval.cpu().numpy(), nz=nz.cpu().numpy(),
                                       rhs=rhs.cpu().numpy(), S=S.cpu().numpy(), g=g.cpu().numpy()))
    # collective 2 + backward step
    lam_own = torch.arange(L.col_begin, L.col_end, dtype=torch.float64, device=dev) * 1e-3
    lam = asm.allgather_lambda(lam_own)              # ghb_allgather_lambda_f64: grouped ncclBroadcast, unequal ranges
    u = torch.empty((n, plan.n_i), dtype=torch.float64, device=dev)
    ctx.backsub(plan, n, A, b, lam, dv, asm.cell_ids[:n].contiguous(), u, info)
    ok = True
    if rank == 0:
        Sg = np.concatenate([p["S"] for p in parts]); gg = np.concatenate([p["g"] for p in parts])
        ctx0 = gh.Context(local)
        ng = L0.ncells_global
        nnz = ctx0.assemble_symbolic(ng, 36, ids_all, L0.nrows_global)
        cp0 = torch.empty(L0.nrows_global + 1, dtype=torch.int64, device=dev); rv0 = torch.empty(nnz, dtype=torch.int64, device=dev)
        ctx0.assemble_pattern(cp0, rv0)
        z0 = torch.empty(nnz, dtype=torch.float64, device=dev); r0 = torch.empty(L0.nrows_global, dtype=torch.float64, device=dev)
        ctx0.assemble_numeric(torch.as_tensor(Sg, device=dev), torch.as_tensor(gg, device=dev), dv, z0, r0)
        off, cat = 0, [np.array([1], dtype=np.int64)]
        for p in parts:
            cat.append(p["colptr"][1:] + off); off += p["colptr"][-1] - 1
        ok &= np.array_equal(np.concatenate(cat), cp0.cpu().numpy())
        ok &= np.array_equal(np.concatenate([p["rowval"] for p in parts]), rv0.cpu().numpy())
        ok &= np.array_equal(np.concatenate([p["nz"] for p in parts]), z0.cpu().numpy())
        ok &= np.array_equal(np.concatenate([p["rhs"] for p in parts]), r0.cpu().numpy())
        ok &= bool(torch.equal(lam, torch.arange(1, L0.nrows_global + 1, dtype=torch.float64, device=dev) * 1e-3))
        # backward step of rank 0's slab against the oracle
        from oracle import oracle as o
        from oracle import oracle_c as oc
        op = o.BlockPlan(ndofs, touched, [1, 2], [3])
        xk = o.cell_dof_values(lam.cpu().numpy(), dv.cpu().numpy(), asm.cell_ids[:n].cpu().numpy())
        u0, _ = oc.backsub(op, A.cpu().numpy(), b.cpu().numpy(), xk)
        err = np.linalg.norm(u.cpu().numpy() - u0, axis=1) / np.linalg.norm(u0, axis=1)
        ok &= bool(err.max() < 1e-11)
        print(f"NCCL_SLAB_CHECK world={world} gdims={gdims} nnz={nnz} ok={bool(ok)} backsub_err={err.max():.2e}")
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    ctx.comm_destroy()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
