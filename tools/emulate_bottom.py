"""Lane-level numpy emulation of the left-looking bottom phase of condense_dmma_ll_kernel (csrc/condense_dmma.cu).

Design aid and index check for the register-resident bottom block: every "register" is a (32,) array over the lanes
of a warp and dmma() reproduces the fragment layout of mma.sync.m8n8k4.f64 (A: lane(gid,tig) holds A[gid][tig],
B: B[tig][gid], C/D: D[gid][2tig], D[gid][2tig+1]).  The accumulator fragment of one product is fed back as the A
operand of the next one without a shared-memory round trip: k-step s of lane tig stands for logical column
SIGMA[2 tig + s], and the B operands are loaded with the matching row permutation.
Run: python tools/emulate_bottom.py  (no GPU, no oracle import; checks against numpy.linalg)."""
import numpy as np

SIGMA = [0, 2, 1, 3, 6, 4, 7, 5]          # logical column (inside a tile) of accumulator column n
LANE = np.arange(32)
GID, TIG = LANE >> 2, LANE & 3


def pc(c):
    return c ^ ((c >> 2) & 1)


def dmma(d0, d1, a, b):
    A = np.zeros((8, 4)); B = np.zeros((4, 8))
    A[GID, TIG] = a
    B[TIG, GID] = b
    D = A @ B
    return d0 + D[GID, 2 * TIG], d1 + D[GID, 2 * TIG + 1]


def bank_check(addr_doubles, lanes_per_phase=16):
    """8-byte accesses: a phase (half-warp) is conflict free iff its 8-byte bank pairs (addr mod 16) are distinct
    or equal addresses."""
    worst = 1
    for h in range(0, 32, lanes_per_phase):
        a = addr_doubles[h:h + lanes_per_phase]
        banks = {}
        for x in a:
            banks.setdefault(x % 16, set()).add(x)
        worst = max(worst, max(len(v) for v in banks.values()))
    return worst


def bottom_tile(Wt_mem, LDW, Dinv, NI, NB, Bt_rows, I, conflicts):
    """One row tile I of the bottom block.  Wt_mem: flat shared-memory image (column-major, leading dimension LDW,
    in-tile column permutation pc) after the top phase; Dinv[p]: flat 64 (k + 8 n); Bt_rows: dense [NB][N+1] of
    [A21 A22 b2] (what the lanes read from the record).  Returns dict (row, col) -> value of S|g for the tile."""
    N = NI + NB
    NC = N + 1
    CT = (NC + 7) // 8
    NP = (NI + 7) // 8
    r = 8 * I + GID
    rv = r < NB
    rho = [np.array([SIGMA[2 * t + s] for t in TIG]) for s in (0, 1)]
    sg = np.array([SIGMA[g] for g in GID])
    x = np.zeros((CT, 2, 32))
    for J in range(CT):
        for e in range(2):
            c = 8 * J + np.array([SIGMA[2 * t + e] for t in TIG])
            ok = rv & (c <= N)
            x[J, e] = np.where(ok, Bt_rows[np.minimum(r, NB - 1), np.minimum(c, N)], 0.0)
    for p in range(NP):
        c0 = 8 * p
        npiv = min(8, NI - c0)
        y0 = np.zeros(32); y1 = np.zeros(32)
        for s in range(2):
            if min(rho[s]) >= npiv:
                continue
            bd = -Dinv[p][rho[s] + 8 * sg]
            y0, y1 = dmma(y0, y1, x[p, s], bd)
        if npiv < 8:
            for s in range(2):
                if min(rho[s]) >= npiv:
                    continue
                ok = (rho[s] < npiv) & (sg >= npiv)
                addr = c0 + rho[s] + LDW * (c0 + np.array([pc(v) for v in sg]))
                ub = np.where(ok, Wt_mem[np.where(ok, addr, 0)], 0.0)
                x[p, 0], x[p, 1] = dmma(x[p, 0], x[p, 1], (y0, y1)[s], ub)
        for J in range(p + 1, CT):
            for s in range(2):
                if min(rho[s]) >= npiv:
                    continue
                ok = rho[s] < npiv
                addr = c0 + rho[s] + LDW * (8 * J + np.array([pc(v) for v in sg]))
                conflicts.append(bank_check(addr))
                ub = np.where(ok, Wt_mem[np.where(ok, addr, 0)], 0.0)
                x[J, 0], x[J, 1] = dmma(x[J, 0], x[J, 1], (y0, y1)[s], ub)
    out = {}
    for J in range(CT):
        for e in range(2):
            c = 8 * J + np.array([SIGMA[2 * t + e] for t in TIG])
            for l in range(32):
                if rv[l] and NI <= c[l] <= N:
                    out[(r[l], c[l] - NI)] = x[J, e, l]
    return out


def run(NI, NB, seed=0):
    rng = np.random.default_rng(seed)
    N = NI + NB
    A = rng.standard_normal((N, N + 1))
    A[:NI, :NI] += 2 * np.sqrt(NI) * np.eye(NI)           # no pivoting needed: the top phase is not under test
    # top phase by numpy: LU without pivoting, W12 = L^-1 [A12 b1]
    T = A[:NI].copy()
    for k in range(NI):
        T[k + 1:, k] /= T[k, k]
        T[k + 1:, k + 1:] -= np.outer(T[k + 1:, k], T[k, k + 1:])
    LDW = ((NI + 3) // 8) * 8 + 4
    if LDW < NI:
        LDW += 8
    CT = (N + 1 + 7) // 8
    Wt = np.full(CT * 8 * LDW + 64, np.nan)
    for c in range(N + 1):
        mc = 8 * (c // 8) + pc(c % 8)
        Wt[LDW * mc: LDW * mc + NI] = T[:, c]
    NP = (NI + 7) // 8
    Dinv = []
    for p in range(NP):
        c0 = 8 * p
        npiv = min(8, NI - c0)
        D = np.zeros((8, 8))
        D[:npiv, :npiv] = np.linalg.inv(np.triu(T[c0:c0 + npiv, c0:c0 + npiv]))
        Dinv.append(D.flatten(order="F"))                 # k + 8 n
    conflicts = []
    got = np.full((NB, NB + 1), np.nan)
    for I in range((NB + 7) // 8):
        for (rr, cc), v in bottom_tile(Wt, LDW, Dinv, NI, NB, A[NI:], I, conflicts).items():
            got[rr, cc] = v
    ref = A[NI:, NI:] - A[NI:, :NI] @ np.linalg.solve(A[:NI, :NI], A[:NI, NI:])
    err = np.abs(got - ref).max() / np.abs(ref).max()
    print(f"({NI},{NB}): max rel err {err:.2e}, worst B-fragment bank conflict degree {max(conflicts)}")
    assert err < 1e-12 and not np.isnan(got).any()
    assert max(conflicts) == 1, "B-fragment loads must be bank-conflict free"


if __name__ == "__main__":
    for shape in [(34, 36), (33, 12), (56, 16)]:
        run(*shape)
