"""world_size-2 (and 3) gloo tests on CPU of the host-side multi-GPU logic: slab layout, cut-plane
exchange plumbing and the lambda all-gather.  (Kernels themselves are covered by the gpu tests.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as o


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, gdims, nf, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gridaphybrid_b200.distributed import SlabLayout, exchange_cut_plane, halo_lambda
        L = SlabLayout(gdims, nf, rank, world)
        ids = L.slab_cell_ids("cpu")
        # every rank's ids must be the rows of the global table (own cells + ghost layer)
        cwf = o.cartesian_cell_wise_facets(gdims)
        fids, nfree, _ = o.facet_dof_ids(o.facet_is_boundary(cwf), nf)
        ref = o.restrict_facet_dofs_to_skeleton(cwf, fids)
        ok = np.array_equal(ids.numpy(), ref[L.cell_start:L.cell_start + L.ncells + L.nghost])
        ok = ok and L.nrows_global == nfree
        # owned columns are exactly the free dofs first touched by own cells
        own = ref[L.cell_start:L.cell_start + L.ncells]
        prev = ref[:L.cell_start]
        prevmax = int(prev[prev > 0].max()) if (prev > 0).any() else 0
        ok = ok and L.col_begin == prevmax + 1 and L.col_end == int(own[own > 0].max()) + 1
        # ghost cells touch owned columns only through their leading nf local dofs
        if L.nghost:
            gh_ids = ids[L.ncells:]
            inrange = (gh_ids >= L.col_begin) & (gh_ids < L.col_end)
            ok = ok and bool(inrange[:, :nf].all()) and not bool(inrange[:, nf:].any())
        # collective 1: cut-plane exchange (rank r>0 -> r-1)
        stride = L.n_b * nf + nf
        send = torch.full((L.layer, stride), float(rank), dtype=torch.float64) if rank > 0 else None
        recv = torch.empty((L.nghost, stride), dtype=torch.float64) if L.nghost else None
        exchange_cut_plane(send, recv, rank, world)
        if recv is not None:
            ok = ok and bool((recv == float(rank + 1)).all())
        # collective 2: all-gather of the owned lambda ranges == global vector
        lam = torch.arange(L.col_begin, L.col_end, dtype=torch.float64)
        full = halo_lambda(lam, L)
        ok = ok and torch.equal(full, torch.arange(1, nfree + 1, dtype=torch.float64))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("gdims,world", [((3, 2, 4), 2), ((4, 6), 2), ((2, 2, 6), 3)])
def test_slab_layout_and_exchange_gloo(gdims, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, gdims, 2, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]
