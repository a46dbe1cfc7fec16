"""The library bar SURVEY section 2a sets: the same condensation with cuBLAS batched calls on the same box --
cublasDgetrfBatched(A11) + cublasDgetrsBatched([A12 b1]) + cublasDgemmBatched(S = A22 - A21 X) through torch's bindings
(torch.linalg.lu_factor / lu_solve / baddbmm on batches of dense FP64 blocks call exactly these).  TOOLS ONLY: nothing of
this is on the product path.  Inputs are dense pre-split blocks already resident in HBM (the packing cost of the
library path is not charged); timing with CUDA events on inputs larger than L2.
usage: python tools/cublas_bar.py [ncells_log2=18]"""
import sys
import torch

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 18
n = 1 << lg
ev = lambda: torch.cuda.Event(enable_timing=True)
print(f"torch {torch.__version__}, {torch.cuda.get_device_name(0)}, {n} cells per batch")
for name, ni, nb in [("C3 (34,36)", 34, 36), ("C2 k=2 (33,12)", 33, 12), ("(40,36)", 40, 36), ("(21,16)", 21, 16), ("C2 k=3 (56,16)", 56, 16)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    A11 = torch.randn(n, ni, ni, dtype=torch.float64, device="cuda", generator=g) + 2 * ni ** 0.5 * torch.eye(ni, dtype=torch.float64, device="cuda")
    R12 = torch.randn(n, ni, nb + 1, dtype=torch.float64, device="cuda", generator=g)       # [A12 | b1]
    A21 = torch.randn(n, nb, ni, dtype=torch.float64, device="cuda", generator=g)
    R22 = torch.randn(n, nb, nb + 1, dtype=torch.float64, device="cuda", generator=g)       # [A22 | b2]

    def step():
        LU, piv = torch.linalg.lu_factor(A11)            # cublasDgetrfBatched
        X = torch.linalg.lu_solve(LU, piv, R12)          # cublasDgetrsBatched
        return torch.baddbmm(R22, A21, X, alpha=-1.0)    # cublasDgemmBatched / strided batched: [S | g]

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    reps = 5
    e0.record()
    for _ in range(reps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # phase split
    ts = []
    for f in (lambda: torch.linalg.lu_factor(A11), None, None):
        pass
    LU, piv = torch.linalg.lu_factor(A11)
    X = torch.linalg.lu_solve(LU, piv, R12)
    parts = []
    for fn in (lambda: torch.linalg.lu_factor(A11), lambda: torch.linalg.lu_solve(LU, piv, R12), lambda: torch.baddbmm(R22, A21, X, alpha=-1.0)):
        fn(); torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        for _ in range(3):
            fn()
        b.record(); torch.cuda.synchronize()
        parts.append(a.elapsed_time(b) / 3)
    print(f"{name:16s} cuBLAS batched getrf+getrs+gemm: {n / ms / 1e3:8.2f} M cells/s  ({ms:.2f} ms: getrf {parts[0]:.2f}, getrs {parts[1]:.2f}, gemm {parts[2]:.2f})", flush=True)
    del A11, R12, A21, R22, LU, piv, X, out
