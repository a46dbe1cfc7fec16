#!/usr/bin/env python
"""bench.py -- cells/s condensed+assembled (FP64) on the 3-D HDG k=2 configuration (BASELINE.json C3).

A "step" is one pass of the hot path over one batch of synthetic cells: static condensation of every
cell (A_K,b_K)->(S_K,g_K) followed by the numeric skeleton assembly into the cached CSC pattern.
Inputs are resident in HBM before the timed region (`value`); `e2e` repeats the same metric through the
C-ABI call `ghb_condense_assemble_f64` with pinned HOST buffers (H2D of the records and D2H of
nzval/rhs inside the timed region) on a bounded sample of the same workload.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dims nx ny nz]

Multi-GPU: one process per GPU (torchrun); cells are partitioned in z-slabs, weak scaling (each rank
owns a dims-sized slab of a mesh that is N times taller).  Timing: CUDA events, barrier +
synchronize on both sides, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NDOFS = [30, 4, 36]          # C3: u in [P2]^3 (30), p in P1 (4), lambda in P2(facet) 6 x 6 (36)
INTERIOR, BOUNDARY = [1, 2], [3]
N_I, N_B = 34, 36
METRIC = "cells/s condensed+assembled (FP64, 3D HDG k=2)"


# dram__bytes_read.sum + dram__bytes_write.sum of condense_dmma_ll_kernel<34,36> from the committed `ncu --set full`
# capture, per cell (5.4339 GB + 1.3871 GB for a 131 072-cell launch); algorithmic bytes are 50 416 B/cell
NCU_DRAM_BYTES_PER_CELL = 52040
NCU_TRAFFIC_SOURCE = "profiles/r01_condense_dmma_summary.md (ncu --set full, 131072-cell launch, scaled per cell)"


def algorithmic_bytes_per_cell():
    """B_cond = 8[(n^2+n) + n_b^2 + n_b]  (read A_K,b_K once, write S_K,g_K once)."""
    n = N_I + N_B
    return 8 * ((n * n + n) + N_B * N_B + N_B)


def flops_per_cell():
    return (2.0 / 3.0) * N_I ** 3 + 2 * N_I ** 2 * (N_B + 1) + 2 * N_I * N_B * (N_B + 1)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.stop, self.t = index, [], threading.Event(), None

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(float(r[2]) for r in self.rows)}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle's C twin (condensation) + SciPy COO->CSC (assembly) on a bounded sample
# --------------------------------------------------------------------------------------------------

def cpu_sample(sample_dims, steps, warmup):
    from oracle import oracle as o
    from oracle import oracle_c as oc
    import scipy.sparse as sp
    plan = o.BlockPlan(NDOFS, np.ones((3, 3), bool), INTERIOR, BOUNDARY)
    n = int(np.prod(sample_dims))
    rng = np.random.default_rng(20261017)
    A = rng.uniform(-1, 1, (n, plan.lenA))
    A[:, :30 * 30:31] += 7.0     # keep A11 comfortably non-singular (values do not affect the timing)
    b = rng.uniform(-1, 1, (n, plan.lenb))
    cwf = cartesian_cwf_numpy(sample_dims)
    ids = facet_ids_numpy(cwf, sample_dims, 6)
    nfree = int(ids.max())
    cores = oc.max_threads()
    li = np.tile(np.arange(N_B), N_B)
    lj = np.repeat(np.arange(N_B), N_B)

    def step():
        S, g, info = oc.condense(plan, A, b, nthreads=0)
        I = ids[:, li].ravel(); J = ids[:, lj].ravel()
        keep = (I > 0) & (J > 0)
        M = sp.coo_matrix((S.ravel()[keep], (I[keep] - 1, J[keep] - 1)), shape=(nfree, nfree)).tocsc()
        rhs = np.zeros(nfree)
        m = ids > 0
        np.add.at(rhs, ids[m] - 1, g[m])
        return M, rhs

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return n / dt, cores, dt, f"{n} cells ({'x'.join(map(str, sample_dims))} mesh): C oracle (pthreads, {cores} threads) " \
                              f"condensation + SciPy coo->csc assembly; Julia reference not runnable here (JULIA_NUM_THREADS=n/a)"


def cartesian_cwf_numpy(dims):
    """closed-form first-touch facet ids (vectorised twin of the device kernel; bench CPU arm only)."""
    D = len(dims)
    stride = np.concatenate([[1], np.cumprod(dims)])
    c = np.arange(stride[-1], dtype=np.int64)

    def own(cc, axis, side):
        run = D * cc
        for a in range(D):
            hi, rem = cc // stride[a + 1], cc % stride[a + 1]
            run = run + hi * stride[a] + np.minimum(rem, stride[a])
        res = np.zeros_like(cc)
        for a in range(D - 1, -1, -1):
            ia = (cc // stride[a]) % dims[a]
            run = run + (ia == 0)
            if a == axis and side == 0:
                res = run.copy()
            run = run + 1
            if a == axis and side == 1:
                res = run.copy()
        return res

    out = np.zeros((len(c), 2 * D), dtype=np.int64)
    for lf in range(2 * D):
        axis, side = D - 1 - lf // 2, lf % 2
        ia = (c // stride[axis]) % dims[axis]
        if side == 1:
            out[:, lf] = own(c, axis, 1)
        else:
            lowb = own(c, axis, 0)
            nb = own(np.maximum(c - stride[axis], 0), axis, 1)
            out[:, lf] = np.where(ia == 0, lowb, nb)
    return out


def facet_ids_numpy(cwf, dims, ndofs_f):
    nfacets = int(cwf.max())
    cnt = np.bincount(cwf.ravel() - 1, minlength=nfacets)
    isd = cnt == 1
    free_rank = np.cumsum(~isd)
    dir_rank = np.cumsum(isd)
    d = np.arange(1, ndofs_f + 1)[None, :]
    fid = np.where(isd[:, None], -((dir_rank[:, None] - 1) * ndofs_f + d), (free_rank[:, None] - 1) * ndofs_f + d)
    return fid[cwf - 1].reshape(cwf.shape[0], -1)


def workload_name(dims):
    """the one workload both arms report (the reference arm times a bounded sample of it)"""
    return (f"C3 Darcy HDG k=2 3-D hex Cartesian {dims[0]}x{dims[1]}x{dims[2]} cells per GPU "
            f"(n_i=34, n_b=36), all boundary facets Dirichlet, Philox synthetic records")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_dims = (32, 32, 24)
    val, cores, dt, sample = cpu_sample(sample_dims, args.steps, max(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "cells/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(tuple(args.dims)), "cells_per_gpu": int(np.prod(args.dims)),
                       "sample": sample},
            "cpu_baseline": {"value": val, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist
    import gridaphybrid_b200 as gh

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N>1)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = gh.Context(local)
    gh.set_default_context(ctx)
    ctx.use_torch_stream()
    dev = torch.device("cuda", local)

    dims = tuple(args.dims)
    gdims = (dims[0], dims[1], dims[2] * world)      # weak scaling: every rank owns a dims-sized z-slab
    ncells = int(np.prod(dims))
    cell_start = rank * ncells
    plan = ctx.plan_blocks(NDOFS, np.ones((3, 3), bool), INTERIOR, BOUNDARY)
    from gridaphybrid_b200.distributed import SlabAssembler
    slab = SlabAssembler(ctx, gdims, 6, rank, world)
    A = torch.empty((ncells, plan.lenA), dtype=torch.float64, device=dev)
    b = torch.empty((ncells, plan.lenb), dtype=torch.float64, device=dev)
    ctx.synth_fill(plan, cell_start, ncells, A, b)
    S = torch.empty((ncells, N_B * N_B), dtype=torch.float64, device=dev)
    g = torch.empty((ncells, N_B), dtype=torch.float64, device=dev)
    info = torch.empty((ncells,), dtype=torch.int32, device=dev)
    nzval = torch.empty(slab.nnz, dtype=torch.float64, device=dev)
    rhs = torch.empty(slab.nrows_local, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)
    cond_ms = []

    def step(timed):
        e0, e1 = ev(), ev()
        e0.record()
        ctx.condense(plan, ncells, A, b, S, g, info)
        e1.record()
        slab.assemble(S, g, nzval, rhs)
        if timed:
            cond_ms.append((e0, e1))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    barrier()
    l0 = ctx.launch_count
    with ClockSampler(local) as clk:
        t0, t1 = ev(), ev()
        t0.record()
        for _ in range(args.steps):
            step(True)
        t1.record()
        barrier()
    launches = ctx.launch_count - l0
    ms = t0.elapsed_time(t1) / args.steps
    if world > 1:
        tms = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    assert int(info.abs().sum().item()) == 0, "a cell failed to factorise"
    kernel_ms = float(np.mean([a.elapsed_time(bb) for a, bb in cond_ms]))
    value = ncells * world / (ms * 1e-3)

    # ---- backward map on the same records (reported on its own; outside the timed region) ----
    L = slab.layout
    lam = torch.randn(L.nrows_global, dtype=torch.float64, device=dev)       # stands for the all-gathered lambda
    u = torch.empty((ncells, plan.n_i), dtype=torch.float64, device=dev)
    ids_local = slab.cell_ids[:ncells]
    for _ in range(2):
        ctx.backsub(plan, ncells, A, b, lam, None, ids_local, u, info)
    barrier()
    b0, b1 = ev(), ev()
    b0.record()
    for _ in range(3):
        ctx.backsub(plan, ncells, A, b, lam, None, ids_local, u, info)
    b1.record()
    barrier()
    back_ms = b0.elapsed_time(b1) / 3
    if world > 1:
        tms = torch.tensor([back_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        back_ms = float(tms.item())
    backsub = {"value": ncells * world / (back_ms * 1e-3), "unit": "cells/s", "ms": back_ms,
               "note": "BackwardStaticCondensationMap on the same records (LU recomputed, as the reference does)"}
    del lam, u

    # ---- e2e: C-ABI with pinned host buffers, bounded sample, copies inside the timed region ----
    e2e = None
    cpu_base = None
    if rank == 0 or world > 1:
        sdims = tuple(args.e2e_dims)
        sn = int(np.prod(sdims))
        ssk = gh.CartesianSkeleton(sdims, ctx)
        sM = gh.FacetFESpace(ssk, 6, ssk.facet_is_boundary())
        sass = gh.SparseMatrixAssembler(sM)
        _, _, snnz = sass.symbolic()
        hA = torch.empty((sn, plan.lenA), dtype=torch.float64).pin_memory()
        hb = torch.empty((sn, plan.lenb), dtype=torch.float64).pin_memory()
        hA.copy_(A[:sn]); hb.copy_(b[:sn])
        hz = torch.empty(snnz, dtype=torch.float64).pin_memory()
        hr = torch.empty(sass.nrows, dtype=torch.float64).pin_memory()
        hinfo = torch.empty(sn, dtype=torch.int32).pin_memory()
        for _ in range(2):
            ctx.condense_assemble(plan, sn, hA, hb, None, hz, hr, hinfo)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ctx.condense_assemble(plan, sn, hA, hb, None, hz, hr, hinfo)   # returns after the D2H completed
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.steps
        if world > 1:
            tdt = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
            dt = float(tdt.item())
        e2e = {"value": sn * world / dt, "unit": "cells/s",
               "h2d_bytes_per_step": int(sn * (plan.lenA + plan.lenb) * 8),
               "d2h_bytes_per_step": int((snnz + sass.nrows) * 8 + sn * 4),
               "sample": f"{sn} cells per GPU ({'x'.join(map(str, sdims))}) via ghb_condense_assemble_f64, pinned host buffers"}
        del hA, hb, hz, hr

    # ---- informational: the same step with the records generated on the device from an affine family:
    # expand -> condense -> assemble from per-cell coefficient vectors; nothing of the size of the records crosses PCIe.
    # (Overwrites A, b: last leg.  The `e2e` key above stays the host-record path the contract asks for.)
    devgen = None
    if world == 1:
        ntab = 1 + 2 * 3
        rng = np.random.default_rng(3)
        TA = np.concatenate([A[:1].cpu().numpy(), 1e-3 * rng.standard_normal((ntab - 1, plan.lenA))])
        Tb = np.concatenate([b[:1].cpu().numpy(), 1e-3 * rng.standard_normal((ntab - 1, plan.lenb))])
        fam = gh.AffineRecordFamily(TA, Tb)
        coef = gh.cartesian_coefficients(dims, tuple(1.0 / d for d in dims), dev)

        def gen_step():
            fam.expand(ctx, plan, coef, A, b)
            ctx.condense(plan, ncells, A, b, S, g, info)
            slab.assemble(S, g, nzval, rhs)

        gen_step()
        barrier()
        g0, g1 = ev(), ev()
        g0.record()
        for _ in range(3):
            gen_step()
        g1.record()
        barrier()
        gms = g0.elapsed_time(g1) / 3
        assert int(info.abs().sum().item()) == 0
        devgen = {"value": ncells / (gms * 1e-3), "unit": "cells/s", "ms": gms, "h2d_bytes_per_step": int(coef.numel() * 8),
                  "note": "records of an affine family (7 tables) generated on the device by ghb_expand_records_f64, then "
                          "condensed and assembled; coefficients counted as the host input"}
        del coef

    if rank == 0:
        if world == 1 and not args.no_cpu:
            v, cores, _, sample = cpu_sample((32, 32, 24), 1, 1)
            cpu_base = {"value": v, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample}
        peak, how = measured_peaks()
        ach = algorithmic_bytes_per_cell() * ncells / (kernel_ms * 1e-3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(dims),
                           "cells_per_gpu": ncells, "l2": "inputs larger than L2 (no flush needed)",
                           "kernel": plan.kernel_name, "nnz_per_gpu": int(slab.nnz)},
                "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                             "traffic": NCU_DRAM_BYTES_PER_CELL * ncells, "traffic_source": NCU_TRAFFIC_SOURCE,
                             "kernel": "condense_dmma_ll_kernel<34,36> (" + plan.kernel_name + ")",
                             "kernel_ms": kernel_ms, "peak_source": how,
                             "fp64_tflops": flops_per_cell() * ncells / (kernel_ms * 1e-3) / 1e12},
                "cpu_baseline": cpu_base, "backsub": backsub, "device_generated_records": devgen}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    # keep stdout clean for the ONE JSON line: libraries (NCCL's version banner, ...) write to fd 1 too
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dims", type=int, nargs=3, default=[128, 128, 128])
    ap.add_argument("--e2e-dims", type=int, nargs=3, default=[64, 64, 32])
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
