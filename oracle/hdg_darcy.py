"""Element matrices of the reference's Darcy HDG test problem on Cartesian meshes (numpy).

TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).  Restates the weak form of
/root/reference/test/DarcyHDGTests.jl:125-135 (including the `(n . n_o)` owner-normal factor of the
stabilisation terms) and its manufactured solution (:9-14, u = 1 + x componentwise, p = -3.14,
f = u + grad p, div u = D) so that the composition condense -> assemble -> solve -> back-substitute can
be checked with the reference's own criterion  ||u - u_h||_L2 < 1e-12  (:142).

Spaces as in the test (:35-37): u in [P_k]^D, p in P_{k-1}, lambda in P_k(facet), all `space=:P`
(total degree), L2 conforming.  Bases here are monomials in the cell-/facet-local coordinates
xi in [0,1]^d; the discrete solution as a function does not depend on the basis, only its dof vector.

Cell-local facet order follows Gridap's HEX/QUAD face numbering: axis = D-1-lf//2,
side = lf%2.  The facet owner is its first (lowest id) cell (src/Skeleton.jl:113-145 via FaceToCellGlue).
"""
from __future__ import annotations

import itertools

import numpy as np


def monomials(d, k):
    """Exponent tuples of P_k in d variables, ordered by total degree then lexicographically."""
    out = []
    for deg in range(k + 1):
        for e in itertools.product(range(deg + 1), repeat=d):
            if sum(e) == deg:
                out.append(e)
    return out


def _gauss01(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def _tensor_rule(d, n):
    x, w = _gauss01(n)
    if d == 0:
        return np.zeros((1, 0)), np.ones(1)
    pts = np.array(list(itertools.product(x, repeat=d)))
    wts = np.prod(np.array(list(itertools.product(w, repeat=d))), axis=1)
    return pts, wts


def _eval(expo, xi):
    """monomial values [npts, nbasis]"""
    xi = np.atleast_2d(xi)
    out = np.ones((xi.shape[0], len(expo)))
    for m, e in enumerate(expo):
        for a, p in enumerate(e):
            if p:
                out[:, m] *= xi[:, a] ** p
    return out


def _eval_grad(expo, xi, axis):
    """d/dxi_axis of the monomials [npts, nbasis]"""
    xi = np.atleast_2d(xi)
    out = np.zeros((xi.shape[0], len(expo)))
    for m, e in enumerate(expo):
        if e[axis] == 0:
            continue
        v = np.full(xi.shape[0], float(e[axis]))
        for a, p in enumerate(e):
            q = p - 1 if a == axis else p
            if q:
                v = v * xi[:, a] ** q
        out[:, m] = v
    return out


class DarcyHDG:
    def __init__(self, dims, order=1, tau=1.0, length=1.0):
        self.dims = tuple(int(d) for d in dims)
        self.D = len(self.dims)
        self.k = int(order)
        self.tau = float(tau)
        self.h = np.array([length / d for d in self.dims])
        self.eu = monomials(self.D, self.k)
        self.ep = monomials(self.D, self.k - 1)
        self.el = monomials(self.D - 1, self.k)
        self.Nu, self.Np, self.Nl = len(self.eu), len(self.ep), len(self.el)
        self.nlf = 2 * self.D
        self.ndofs = [self.D * self.Nu, self.Np, self.nlf * self.Nl]
        self.ncells = int(np.prod(self.dims))
        self.nq = self.k + 2

    # ---- geometry ---------------------------------------------------------------------------
    def cell_index(self, c):
        idx = []
        for d in range(self.D):
            idx.append(c % self.dims[d])
            c //= self.dims[d]
        return idx

    def lfacet(self, lf):
        return self.D - 1 - lf // 2, lf % 2   # (axis, side)

    def n_dot_no(self, idx, lf):
        axis, side = self.lfacet(lf)
        if side == 1 or idx[axis] == 0:
            return 1.0     # high-side facet (this cell is the first cell around it) or boundary facet
        return -1.0        # low-side interior facet: owned by the lower neighbour

    # ---- exact solution (test/DarcyHDGTests.jl:9-14) ----------------------------------------------
    @staticmethod
    def p_exact():
        return -3.14

    def u_exact_coeffs(self, idx):
        """dofs of u = (1+x_d)_d on the cell in the monomial basis: 1 + x0_d + h_d*xi_d."""
        c = np.zeros(self.D * self.Nu)
        for d in range(self.D):
            x0 = idx[d] * self.h[d]
            c[d * self.Nu + self.eu.index(tuple(0 for _ in range(self.D)))] = 1.0 + x0
            e = tuple(1 if a == d else 0 for a in range(self.D))
            c[d * self.Nu + self.eu.index(e)] = self.h[d]
        return c

    # ---- element matrices -----------------------------------------------------------------------
    def cell_blocks(self):
        """Returns (mats, vecs, touched): mats[i][j] float64 [ncells, r_i, c_j], vecs[i] [ncells, r_i]
        for the fields (u, p, lambda); all 9 blocks touched."""
        D, Nu, Np, Nl, nlf = self.D, self.Nu, self.Np, self.Nl, self.nlf
        h = self.h
        vol = float(np.prod(h))
        xq, wq = _tensor_rule(D, self.nq)
        U = _eval(self.eu, xq)
        P = _eval(self.ep, xq)
        Muu = (U * wq[:, None]).T @ U * vol
        Auu = np.zeros((D * Nu, D * Nu))
        Aup = np.zeros((D * Nu, Np))
        Apu_vol = np.zeros((Np, D * Nu))
        for d in range(D):
            Auu[d * Nu:(d + 1) * Nu, d * Nu:(d + 1) * Nu] = Muu
            dU = _eval_grad(self.eu, xq, d) / h[d]
            dP = _eval_grad(self.ep, xq, d) / h[d]
            Aup[d * Nu:(d + 1) * Nu, :] = -(dU * wq[:, None]).T @ P * vol
            Apu_vol[:, d * Nu:(d + 1) * Nu] = -(dP * wq[:, None]).T @ U * vol
        # facet quantities (identical for every cell up to the owner sign)
        xf, wf = _tensor_rule(D - 1, self.nq)
        fac = []
        for lf in range(nlf):
            axis, side = self.lfacet(lf)
            area = float(np.prod([h[a] for a in range(D) if a != axis]))
            xi = np.insert(xf, axis, float(side), axis=1)
            Uf = _eval(self.eu, xi)
            Pf = _eval(self.ep, xi)
            Lf = _eval(self.el, xf)
            w = wf * area
            nsign = 1.0 if side == 1 else -1.0
            fac.append(dict(axis=axis, nsign=nsign,
                            UL=(Uf * w[:, None]).T @ Lf, PU=(Pf * w[:, None]).T @ Uf,
                            PP=(Pf * w[:, None]).T @ Pf, PL=(Pf * w[:, None]).T @ Lf,
                            LL=(Lf * w[:, None]).T @ Lf))
        nc = self.ncells
        nl = nlf * Nl
        mats = [[np.zeros((nc, self.ndofs[i], self.ndofs[j])) for j in range(3)] for i in range(3)]
        vecs = [np.zeros((nc, self.ndofs[i])) for i in range(3)]
        for c in range(nc):
            idx = self.cell_index(c)
            Apu = Apu_vol.copy()
            App = np.zeros((Np, Np))
            Aul = np.zeros((D * Nu, nl))
            Apl = np.zeros((Np, nl))
            All = np.zeros((nl, nl))
            for lf, F in enumerate(fac):
                a, ns = F["axis"], F["nsign"]
                s = self.n_dot_no(idx, lf)
                sl = slice(lf * Nl, (lf + 1) * Nl)
                Aul[a * Nu:(a + 1) * Nu, sl] += ns * F["UL"]           # (vh.n)*lh
                Apu[:, a * Nu:(a + 1) * Nu] += ns * F["PU"]            # qh*(uh.n)
                App += self.tau * s * F["PP"]                          # tau*qh*ph*(n.no)
                Apl[:, sl] += -self.tau * s * F["PL"]                  # -tau*qh*lh*(n.no)
                All[sl, sl] += -self.tau * s * F["LL"]                 # -tau*mh*lh*(n.no)
            mats[0][0][c] = Auu
            mats[0][1][c] = Aup
            mats[0][2][c] = Aul
            mats[1][0][c] = Apu
            mats[1][1][c] = App
            mats[1][2][c] = Apl
            mats[2][0][c] = Aul.T                                      # mh*(uh.n)
            mats[2][1][c] = -Apl.T                                     # tau*mh*ph*(n.no)
            mats[2][2][c] = All
            # l(v,q,m) = int v.f + q*(div u),  f = u_exact, div u = D
            x = np.array([(idx[d] + xq[:, d]) * h[d] for d in range(D)]).T
            for d in range(D):
                vecs[0][c, d * Nu:(d + 1) * Nu] = (U * (wq * (1.0 + x[:, d]))[:, None]).sum(0) * vol
            vecs[1][c] = (P * (wq * float(D))[:, None]).sum(0) * vol
        touched = np.ones((3, 3), dtype=bool)
        return mats, vecs, touched

    def dirichlet_values(self, ndir_facets):
        """lambda = p on Dirichlet facets: constant => coefficient of the constant monomial."""
        v = np.zeros((ndir_facets, self.Nl))
        v[:, 0] = self.p_exact()
        return v.reshape(-1)

    def l2_error_u(self, u_cells):
        """sqrt(sum_K (c - c_ex)^T M (c - c_ex)) with u_cells [ncells, D*Nu]."""
        xq, wq = _tensor_rule(self.D, self.nq)
        U = _eval(self.eu, xq)
        M = (U * wq[:, None]).T @ U * float(np.prod(self.h))
        err2 = 0.0
        for c in range(self.ncells):
            dcf = np.asarray(u_cells[c]) - self.u_exact_coeffs(self.cell_index(c))
            for d in range(self.D):
                e = dcf[d * self.Nu:(d + 1) * self.Nu]
                err2 += float(e @ M @ e)
        return float(np.sqrt(err2))
