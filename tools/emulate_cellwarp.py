"""Lane-level numpy emulation of condense_cw_kernel (csrc/condense_cw.cu): one warp = one cell, no barriers.

Executable specification of the index arithmetic: every "register" is a (32,) array over the lanes of a warp, dmma()
reproduces the fragment layout of mma.sync.m8n8k4.f64 (A: lane (g,t) holds A[g][t]; B: B[t][g]; C/D: D[g][2t], D[g][2t+1])
and shared memory is an explicit array.  All register tiles are held TRANSPOSED (tile[g][c] = W[row slot c][column g]):
a D fragment is then directly the A operand of the next product (k-step e of lane t stands for column 2t+e), so the
triangular solves and the Schur update chain in registers and only the B operands come from shared / global memory.

  LU phase   left-looking by column tiles of A11; rows never move in shared memory (row-major image by ORIGINAL row,
             leading dimension 40, 16-byte chunks XOR-swizzled by (row>>1)&3); perm[slot] = original row at that
             position.  Panel factorisation with one row per lane (implicit pivoting, two register sets for the first
             panel only: the rows 32.. move into retired lanes afterwards).  Multipliers and the off-diagonal U tiles
             are stored negated; inv(L_pp), inv(U_pp) as 8x8 tiles [n][k].
  per column tile J of [A12 b1]:  Z = L^-1 P A12_J,  X = U^-1 Z,  S_J = A22_J - A21 X  (all as transposed tiles).

Run: python tools/emulate_cellwarp.py   (no GPU, no oracle import; checks against numpy.linalg and reports the
shared-memory wavefront counts of the access patterns under the swizzle)."""
import numpy as np

LANE = np.arange(32)
G, T = LANE >> 2, LANE & 3
LDL = 40


def dmma(d, a, b):
    A = np.zeros((8, 4)); B = np.zeros((4, 8))
    A[G, T] = a
    B[T, G] = b
    D = A @ B
    return [d[0] + D[G, 2 * T], d[1] + D[G, 2 * T + 1]]


class Smem:
    """Row-major image [rows][LDL] with swizzled 16-byte chunks; counts wavefronts of every access."""

    def __init__(self, rows):
        self.rows = rows
        self.mem = np.zeros(rows * LDL)
        self.wavefronts = 0
        self.ideal = 0

    @staticmethod
    def addr(r, c):      # doubles
        chunk = (c >> 1) ^ ((r >> 1) & 3)
        return r * LDL + 2 * chunk + (c & 1)

    def _count(self, addrs, width, mask):
        """width = doubles per lane (1: 64-bit, 2: 128-bit).  Hardware model: a 64-bit access is served per half-warp,
        a 128-bit access per quarter-warp; conflict degree = max distinct 16-byte... (4-byte bank) rows per bank."""
        lanes_per = 16 if width == 1 else 8
        total = 0
        for h in range(0, 32, lanes_per):
            banks = {}
            for l in range(h, h + lanes_per):
                if not mask[l]:
                    continue
                for w in range(2 * width):
                    word = 2 * addrs[l] + w
                    banks.setdefault(word % 32, set()).add(word // 32)
            total += max((len(v) for v in banks.values()), default=0)
        self.wavefronts += total
        self.ideal += 32 // lanes_per

    def ld64(self, r, c, mask=None):
        mask = np.ones(32, bool) if mask is None else mask
        a = self.addr(r, c)
        self._count(a, 1, mask)
        return np.where(mask, self.mem[np.where(mask, a, 0)], 0.0)

    def ld128(self, r, c, mask=None):   # c even: returns columns c, c+1
        mask = np.ones(32, bool) if mask is None else mask
        a = self.addr(r, c)
        self._count(a, 2, mask)
        a = np.where(mask, a, 0)
        return np.where(mask, self.mem[a], 0.0), np.where(mask, self.mem[a + 1], 0.0)

    def st64(self, r, c, v, mask):
        a = self.addr(r, c)
        self._count(a, 1, mask)
        self.mem[a[mask]] = v[mask]

    def st128(self, r, c, v0, v1, mask):
        a = self.addr(r, c)
        self._count(a, 2, mask)
        self.mem[a[mask]] = v0[mask]
        self.mem[a[mask] + 1] = v1[mask]


def panel_factor(a, act, npiv):
    """One warp, rows in `a` [nset][8][32] (register sets x columns x lanes), act [nset][32] rows still unpivoted.
    Returns ch [nset][32] (step at which the row became pivot, -1 otherwise), rinv[8], info (0 or step+1).
    On return: non-pivot active rows hold the NEGATED multipliers in columns < npiv; pivot row k holds negated
    multipliers in columns < k and its U row in columns >= k."""
    nset = a.shape[0]
    ch = -np.ones((nset, 32), int)
    rinv = np.zeros(8)
    info = 0
    for k in range(npiv):
        cand = act & (ch < 0)
        mag = np.where(cand, np.abs(a[:, k, :]), -1.0)
        # the kernel: REDUX.MAX over (|a| truncated to exponent + 15 mantissa bits, lowest row wins ties)
        s, l = np.unravel_index(np.argmax(mag), mag.shape)
        if mag[s, l] == 0.0 and info == 0:
            info = k + 1
        piv = a[s, k, l]
        rinv[k] = 1.0 / piv if piv != 0 else 0.0
        ch[s, l] = k
        upd = cand.copy(); upd[s, l] = False
        m = np.where(upd, a[:, k, :] * rinv[k], 0.0)
        a[:, k, :] = np.where(upd, -m, a[:, k, :])
        for j in range(k + 1, 8):
            pj = a[s, j, l]                      # SHFL from the pivot lane
            a[:, j, :] = a[:, j, :] - m * pj
    return ch, rinv, info


def run(NI, NB, seed=0, verbose=True):
    rng = np.random.default_rng(seed)
    N = NI + NB
    NC = NB + 1
    RT = (NI + 7) // 8
    CTB = (NC + 7) // 8
    BTM = (NB + 7) // 8
    assert NI <= 64
    DUMMY = NI
    Afull = rng.standard_normal((N, N + 1))
    Afull[:NI, :NI] += 2 * np.sqrt(NI) * np.eye(NI)
    Afull[:NI] = Afull[rng.permutation(NI)]              # forces real pivoting
    A11, A12, A21, A22 = Afull[:NI, :NI], Afull[:NI, NI:], Afull[NI:, :NI], Afull[NI:, NI:]

    W = Smem(NI + 1)
    for r in range(NI):
        for c in range(NI):
            W.mem[W.addr(r, c)] = A11[r, c]
    W.wavefronts = W.ideal = 0
    invL = np.zeros((RT, 64)); invU = np.zeros((RT, 64))
    perm = np.full(8 * RT, DUMMY, int)
    perm[:NI] = np.arange(NI)
    stats = {}

    # lane -> row map of the panel stage: set 0 = row `lane`, set 1 (first panel only) = row 32 + lane
    nset0 = 2 if NI > 32 else 1
    myrow = np.full((nset0, 32), -1, int)
    myrow[0] = np.where(LANE < NI, LANE, -1)
    if nset0 == 2:
        myrow[1] = np.where(32 + LANE < NI, 32 + LANE, -1)
    pivoted = np.zeros((nset0, 32), bool)
    info_cell = 0

    def load_tiles(col0, src_ld, ncols_valid):
        """transposed tiles of an 8-column block: tile[j][e] = M[perm[8j+2t+e]][col0+g]"""
        tiles = []
        for j in range(RT):
            te = []
            for e in range(2):
                r = perm[8 * j + 2 * T + e]
                ok = (r != DUMMY) & (G < ncols_valid)
                te.append(src_ld(r, col0 + G, ok))
            tiles.append(te)
        return tiles

    for R in range(RT):
        c0 = 8 * R
        npiv = min(8, NI - c0)
        w0 = W.wavefronts
        if R > 0:
            Tt = load_tiles(c0, lambda r, c, ok: W.ld64(np.where(ok, r, DUMMY), np.where(ok, c, 0), ok), npiv)
            rowB = [perm[8 * i + G] for i in range(RT)]
            for q in range(R):
                b0, b1 = invL[q][G * 8 + 2 * T], invL[q][G * 8 + 2 * T + 1]
                U = dmma([0.0, 0.0], Tt[q][0], b0)
                U = dmma(U, Tt[q][1], b1)
                Tt[q] = U
                for i in range(q + 1, RT):
                    l0, l1 = W.ld128(rowB[i], np.full(32, 8 * q) + 2 * T)
                    Tt[i] = dmma(Tt[i], U[0], l0)
                    Tt[i] = dmma(Tt[i], U[1], l1)
            for j in range(RT):
                for e in range(2):
                    r = perm[8 * j + 2 * T + e]
                    ok = (r != DUMMY) & (G < npiv)
                    v = -Tt[j][e] if j < R else Tt[j][e]          # off-diagonal U tiles are stored negated
                    W.st64(np.where(ok, r, DUMMY), np.where(ok, c0 + G, 0), v, ok)
        stats.setdefault("lu_update", 0); stats["lu_update"] += W.wavefronts - w0; w0 = W.wavefronts
        # ---- panel: one row per lane
        nset = myrow.shape[0]
        a = np.zeros((nset, 8, 32))
        act = (myrow >= 0) & ~pivoted
        for s in range(nset):
            for c in range(0, 8, 2):
                v0, v1 = W.ld128(np.where(act[s], myrow[s], DUMMY), np.full(32, c0 + c), act[s])
                a[s, c], a[s, c + 1] = v0, v1
        ch, rinv, info = panel_factor(a, act, npiv)
        if info and not info_cell:
            info_cell = c0 + info
        for s in range(nset):
            for c in range(0, 8, 2):
                W.st128(np.where(act[s], myrow[s], DUMMY), np.full(32, c0 + c), a[s, c], a[s, c + 1], act[s])
        # pad columns of a partial panel must stay zero: they were loaded as zero and eliminated with zero pivots rows
        # ---- perm: pivots of this panel, then the rows still in play in lane order
        newp = ch >= 0
        for s in range(nset):
            for l in range(32):
                if newp[s, l]:
                    perm[c0 + ch[s, l]] = myrow[s, l]
        pivoted |= newp
        rest = [(s, l) for s in range(nset) for l in range(32) if myrow[s, l] >= 0 and not pivoted[s, l]]
        perm[c0 + npiv:] = DUMMY
        for rank, (s, l) in enumerate(rest):
            perm[c0 + npiv + rank] = myrow[s, l]
        # ---- inverses of the diagonal block: lanes 0-7 column n of inv(L), lanes 8-15 column n of inv(U)
        D = np.zeros((8, 8))
        for i in range(npiv):
            for c in range(8):
                D[i, c] = W.mem[W.addr(perm[c0 + i], c0 + c)]
        Lm = np.eye(8); Um = np.zeros((8, 8))
        for i in range(npiv):
            for c in range(npiv):
                if c < i:
                    Lm[i, c] = -D[i, c]           # stored negated
                else:
                    Um[i, c] = D[i, c]
        Li = np.zeros((8, 8)); Ui = np.zeros((8, 8))
        # substitution exactly as the lanes do it (column n per lane)
        for n in range(8):
            x = np.zeros(8); x[n] = 1.0 if n < npiv else 0.0
            for i in range(n + 1, npiv):
                x[i] = -sum(Lm[i, m] * x[m] for m in range(n, i))
            Li[:, n] = x
            y = np.zeros(8)
            if n < npiv:
                y[n] = rinv[n]
                for i in range(n - 1, -1, -1):
                    y[i] = -sum(Um[i, m] * y[m] for m in range(i + 1, n + 1)) * rinv[i]
            Ui[:, n] = y
        invL[R] = Li.flatten()        # [n][k] row-major: tile[row*8 + col]
        invU[R] = Ui.flatten()
        stats.setdefault("panel", 0); stats["panel"] += W.wavefronts - w0
        # ---- compaction after the first panel: rows of the second register set move into retired lanes
        if R == 0 and nset == 2:
            free = [l for l in range(32) if pivoted[0, l]]
            nm = myrow[0].copy(); npv = pivoted[0].copy()
            for l in range(32):
                if myrow[1, l] >= 0 and not pivoted[1, l]:
                    d = free.pop(0)
                    nm[d] = myrow[1, l]; npv[d] = False
            myrow = nm[None, :]; pivoted = npv[None, :]

    # ---------------------------------------------------------------- phase B: column tiles of [A12 b1]
    S = np.full((NB, NC), np.nan)
    X = np.full((NI, NC), np.nan)
    rowB = [perm[8 * i + G] for i in range(RT)]
    wB0 = W.wavefronts
    for J in range(CTB):
        nv = min(8, NC - 8 * J)
        Tt = load_tiles(8 * J, lambda r, c, ok: np.where(ok, A12[np.where(ok, r, 0), np.where(ok, c, 0)], 0.0), nv)
        for q in range(RT):
            b0, b1 = invL[q][G * 8 + 2 * T], invL[q][G * 8 + 2 * T + 1]
            Z = dmma([0.0, 0.0], Tt[q][0], b0)
            Z = dmma(Z, Tt[q][1], b1)
            Tt[q] = Z
            for i in range(q + 1, RT):
                l0, l1 = W.ld128(rowB[i], np.full(32, 8 * q) + 2 * T)
                Tt[i] = dmma(Tt[i], Z[0], l0)
                Tt[i] = dmma(Tt[i], Z[1], l1)
        for q in range(RT - 1, -1, -1):
            b0, b1 = invU[q][G * 8 + 2 * T], invU[q][G * 8 + 2 * T + 1]
            Xq = dmma([0.0, 0.0], Tt[q][0], b0)
            Xq = dmma(Xq, Tt[q][1], b1)
            Tt[q] = Xq
            for p in range(q):
                u0, u1 = W.ld128(rowB[p], np.full(32, 8 * q) + 2 * T)
                Tt[p] = dmma(Tt[p], Xq[0], u0)
                Tt[p] = dmma(Tt[p], Xq[1], u1)
        for p in range(RT):
            for e in range(2):
                for l in range(32):
                    k = 8 * p + 2 * T[l] + e
                    if k < NI and G[l] < nv:
                        X[k, 8 * J + G[l]] = Tt[p][e][l]
        # S_J^T = A22_J^T - X_J^T A21^T
        acc = []
        for m in range(BTM):
            te = []
            for e in range(2):
                r = 8 * m + 2 * T + e
                ok = (r < NB) & (G < nv)
                te.append(np.where(ok, A22[np.where(ok, r, 0), np.where(ok, 8 * J + G, 0)], 0.0))
            acc.append(te)
        for p in range(RT):
            for e in range(2):
                k = 8 * p + 2 * T + e
                if (k >= NI).all():
                    continue
                for m in range(BTM):
                    r = 8 * m + G
                    ok = (r < NB) & (k < NI)
                    bf = np.where(ok, A21[np.where(ok, r, 0), np.where(ok, k, 0)], 0.0)
                    acc[m] = dmma(acc[m], -Tt[p][e], bf)
        for m in range(BTM):
            for e in range(2):
                for l in range(32):
                    r = 8 * m + 2 * T[l] + e
                    if r < NB and G[l] < nv:
                        S[r, 8 * J + G[l]] = acc[m][e][l]
    stats["phaseB"] = W.wavefronts - wB0

    Xref = np.linalg.solve(A11, A12)
    Sref = A22 - A21 @ Xref
    eS = np.abs(S - Sref).max() / np.abs(Sref).max()
    eX = np.abs(X - Xref).max() / np.abs(Xref).max()
    if verbose:
        print(f"({NI},{NB}): rel err S {eS:.2e}, X {eX:.2e}; info {info_cell}; smem wavefronts {W.wavefronts} "
              f"(conflict-free {W.ideal}); by phase {stats}")
    assert eS < 1e-12 and eX < 1e-12 and not np.isnan(S).any() and not np.isnan(X).any()
    return W.wavefronts


if __name__ == "__main__":
    for shape in [(34, 36), (33, 12), (40, 36), (21, 16), (45, 24), (56, 16)]:
        run(*shape)
    w = [run(34, 36, seed=s, verbose=False) for s in range(20)]
    print("(34,36) wavefronts over 20 seeds: mean", np.mean(w))
