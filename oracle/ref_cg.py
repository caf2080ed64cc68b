"""CPU restatement of the reference's CG operators (oracle, test infrastructure only).

Follows /root/reference/plancklens/qcinv/opfilt_tt.py (:30-93, :183-205), opfilt_pp.py (:27-107, :253-317),
template_removal.py (:116-150) and cd_solve.py (:35-107, specialised to one preconditioner and tr_cg).
Pinned against the unmodified reference through tests/golden/reference_golden.npz.
"""
import numpy as np

from . import ref_geom as rg
from . import ref_sht as sht
from .healpy_shim.healpy import almxfl


def _cli(cl):
    r = np.zeros_like(cl)
    r[cl != 0] = 1. / cl[cl != 0]
    return r


def _lmax(alm):
    return int(np.floor(np.sqrt(2 * alm.size) - 1))


def dot_tt(a, b):
    """opfilt_tt.py:43-51"""
    lmax = _lmax(a)
    w = np.full(a.size, 2.0)
    w[:lmax + 1] = 1.0
    return float(np.sum(w * (a * np.conj(b)).real))


def dot_pp(a, b):
    """opfilt_pp.py:27-34 (l >= 2 only); a, b = (elm, blm)"""
    lmax = _lmax(a[0])
    ls = np.concatenate([np.arange(m, lmax + 1) for m in range(lmax + 1)])
    w = np.full(a[0].size, 2.0)
    w[:lmax + 1] = 1.0
    w[ls < 2] = 0.0
    return float(np.sum(w * ((a[0] * np.conj(b[0])).real + (a[1] * np.conj(b[1])).real)))


class ninv_tt:
    """opfilt_tt.alm_filter_ninv with monopole + dipole marginalisation (opfilt_tt.py:99-205)."""

    def __init__(self, n_inv, b_transf, marge_monopole=True, marge_dipole=True, marge_maps=()):
        self.n_inv = np.asarray(n_inv, dtype=float)
        self.b = np.asarray(b_transf, dtype=float)
        self.npix = self.n_inv.size
        self.nside = rg.npix2nside(self.npix)
        theta, phi = rg.pix2ang(self.nside)
        modes = [np.asarray(m, dtype=float) for m in marge_maps]     # template maps first (opfilt_tt.py:115-119)
        if marge_monopole:
            modes.append(np.ones(self.npix))
        if marge_dipole:
            modes += [np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)]
        self.P = np.array(modes)
        if len(modes):
            M = np.einsum('ap,p,bp->ab', self.P, self.n_inv, self.P)
            ev, ew = np.linalg.eigh(M)
            self.Minv = ew @ np.diag(1.0 / ev) @ ew.T

    def apply_map(self, t):
        t = t * self.n_inv
        if self.P.shape[0]:
            coeffs = self.Minv @ (self.P @ t)
            t = t - self.n_inv * (coeffs @ self.P)
        return t

    def apply_alm(self, alm):
        lmax = _lmax(alm)
        t = sht.alm2map(almxfl(alm, self.b), self.nside, lmax=lmax)
        t = self.apply_map(t)
        return almxfl(sht.map2alm(t, lmax=lmax), self.b * (self.npix / (4. * np.pi)))

    def calc_prep(self, m):
        lmax = len(self.b) - 1
        return almxfl(sht.map2alm(self.apply_map(np.array(m, dtype=float)), lmax=lmax), self.b * (self.npix / (4. * np.pi)))


def fwd_tt(x, cltt, nf):
    """opfilt_tt.fwd_op.calc (:67-73)"""
    if not np.any(x):
        return x
    return nf.apply_alm(x) + almxfl(x, _cli(cltt))


def pre_diag_tt(cltt, nf):
    """opfilt_tt.pre_op_diag (:76-93)"""
    lmax = len(nf.b) - 1
    f = _cli(cltt[:lmax + 1]) + np.sum(nf.n_inv) / (4.0 * np.pi) * nf.b[:lmax + 1] ** 2
    return _cli(f)


class ninv_pp:
    """opfilt_pp.alm_filter_ninv, 1 or 3 noise maps (opfilt_pp.py:253-303), Q / U template maps (:149-185, :279-290)."""

    def __init__(self, n_inv, b_transf, marge_qmaps=(), marge_umaps=()):
        self.n_inv = [np.asarray(n, dtype=float) for n in n_inv]
        self.tq = [np.asarray(m, dtype=float) for m in marge_qmaps]
        self.tu = [np.asarray(m, dtype=float) for m in marge_umaps]
        self.tniti = []
        for ts in (self.tq, self.tu):
            if len(ts):
                P = np.array(ts)
                ev, ew = np.linalg.eigh(np.einsum('ap,p,bp->ab', P, self.n_inv[0], P))
                self.tniti.append(ew @ np.diag(1.0 / ev) @ ew.T)
            else:
                self.tniti.append(None)
        self.b = np.asarray(b_transf, dtype=float)
        self.npix = self.n_inv[0].size
        self.nside = rg.npix2nside(self.npix)

    def apply_map(self, q, u):
        if len(self.n_inv) == 1:
            q, u = q * self.n_inv[0], u * self.n_inv[0]
            out = []
            for m, ts, ti in ((q, self.tq, self.tniti[0]), (u, self.tu, self.tniti[1])):
                if ti is not None:
                    P = np.array(ts)
                    m = m - self.n_inv[0] * ((ti @ (P @ m)) @ P)
                out.append(m)
            return out[0], out[1]
        return q * self.n_inv[0] + self.n_inv[1] * u, u * self.n_inv[2] + self.n_inv[1] * q

    def apply_alm(self, e, b):
        lmax = _lmax(e)
        q, u = sht.alm2map_spin([almxfl(e, self.b), almxfl(b, self.b)], self.nside, 2, lmax)
        q, u = self.apply_map(q, u)
        te, tb = sht.map2alm_spin([q, u], 2, lmax=lmax)
        f = self.b * (self.npix / (4. * np.pi))
        return almxfl(te, f), almxfl(tb, f)

    def calc_prep(self, q, u):
        lmax = len(self.b) - 1
        q, u = self.apply_map(np.array(q, dtype=float), np.array(u, dtype=float))
        te, tb = sht.map2alm_spin([q, u], 2, lmax=lmax)
        f = self.b * (self.npix / (4. * np.pi))
        return almxfl(te, f), almxfl(tb, f)


def sinv_pp(cls, lmax):
    """opfilt_pp.alm_filter_sinv (:87-107) without EB: diagonal pseudo-inverse"""
    return _cli(cls['ee'][:lmax + 1]), _cli(cls['bb'][:lmax + 1])


def fwd_pp(e, b, cls, nf):
    """opfilt_pp.fwd_op.calc (:51-55)"""
    ne, nb = nf.apply_alm(e, b)
    se, sb = sinv_pp(cls, _lmax(e))
    return ne + almxfl(e, se), nb + almxfl(b, sb)


def pcg(b, fwd, pre, dot, eps_min, iter_max=10000, roundoff=25):
    """cd_solve.cd_solve (:35-107) for one preconditioner with tr_cg, started from x = 0.
    Returns x, iteration count and the (iter, eps) trace of cd_monitors.monitor_basic (:29-41)."""
    add = lambda u, v, a: tuple(ui + a * vi for ui, vi in zip(u, v)) if isinstance(u, tuple) else u + a * v
    zero = tuple(np.zeros_like(c) for c in b) if isinstance(b, tuple) else np.zeros_like(b)
    x = zero
    residual = b            # fwd_op(0) = 0
    d0 = dot(b, b)
    searchdir = pre(residual)
    trace = []
    it = 0
    prev = None
    while True:
        delta = dot(residual, residual)
        trace.append((it, np.sqrt(delta / d0)))
        if it >= iter_max or delta <= eps_min ** 2 * d0:
            break
        searchfwd = fwd(searchdir)
        alpha = dot(searchdir, residual) / dot(searchdir, searchfwd)
        dTAd_inv = 1.0 / dot(searchdir, searchfwd)
        x = add(x, searchdir, alpha)
        prev = (dTAd_inv, searchdir, searchfwd)
        it += 1
        if it % roundoff == 0:
            residual = add(b, fwd(x), -1.0)
        else:
            residual = add(residual, searchfwd, -alpha)
        searchdir = pre(residual)
        beta = prev[0] * dot(searchdir, prev[2])
        searchdir = add(searchdir, prev[1], -beta)
    return x, it, trace


# ---------------------------------------------------------------------------------------------- joint T + P
class ninv_tp:
    """opfilt_tp.alm_filter_ninv (opfilt_tp.py:166-326): n_inv = [TT, PP] or [TT, QQ, QU, UU]; T templates."""

    def __init__(self, n_inv, b_transf, marge_monopole=False, marge_dipole=False, marge_maps_t=()):
        self.t = ninv_tt(n_inv[0], b_transf, marge_monopole=marge_monopole, marge_dipole=marge_dipole, marge_maps=marge_maps_t)
        self.p = ninv_pp(n_inv[1:], b_transf)
        self.n_inv = [np.asarray(n, dtype=float) for n in n_inv]
        self.b = np.asarray(b_transf, dtype=float)
        self.npix, self.nside = self.t.npix, self.t.nside

    def apply_map(self, t, q, u):
        q, u = self.p.apply_map(q, u)
        return self.t.apply_map(t), q, u

    def apply_alm(self, t, e, b):
        lmax = _lmax(t)
        tm = sht.alm2map(almxfl(t, self.b), self.nside, lmax=lmax)
        q, u = sht.alm2map_spin([almxfl(e, self.b), almxfl(b, self.b)], self.nside, 2, lmax)
        tm, q, u = self.apply_map(tm, q, u)
        f = self.b * (self.npix / (4. * np.pi))
        te, tb = sht.map2alm_spin([q, u], 2, lmax=lmax)
        return almxfl(sht.map2alm(tm, lmax=lmax), f), almxfl(te, f), almxfl(tb, f)

    def calc_prep(self, t, q, u):
        lmax = len(self.b) - 1
        tm, q, u = self.apply_map(np.array(t, dtype=float), np.array(q, dtype=float), np.array(u, dtype=float))
        f = self.b * (self.npix / (4. * np.pi))
        te, tb = sht.map2alm_spin([q, u], 2, lmax=lmax)
        return almxfl(sht.map2alm(tm, lmax=lmax), f), almxfl(te, f), almxfl(tb, f)

    def ftebl(self):
        s = lambda m: np.sum(m) / (4.0 * np.pi)
        npp = s(self.n_inv[1]) if len(self.n_inv) == 2 else s(0.5 * (self.n_inv[1] + self.n_inv[3]))
        return s(self.n_inv[0]) * self.b ** 2, npp * self.b ** 2, npp * self.b ** 2


def slinv_tp(cls, lmax):
    """opfilt_tp.alm_filter_sinv (:126-147): per-l pinv of the TEB covariance"""
    m = np.zeros((lmax + 1, 3, 3))
    z = np.zeros(lmax + 1)
    for (i, j), k in {(0, 0): 'tt', (0, 1): 'te', (0, 2): 'tb', (1, 1): 'ee', (1, 2): 'eb', (2, 2): 'bb'}.items():
        m[:, i, j] = m[:, j, i] = cls.get(k, z)[:lmax + 1]
    return np.linalg.pinv(m)


def lmat3(mat, t, e, b):
    v = (t, e, b)
    return tuple(sum(almxfl(v[j], mat[:, i, j]) for j in range(3)) for i in range(3))


def fwd_tp(t, e, b, cls, nf):
    """opfilt_tp.fwd_op.calc (:75-82)"""
    n = nf.apply_alm(t, e, b)
    s = lmat3(slinv_tp(cls, _lmax(t)), t, e, b)
    return tuple(ni + si for ni, si in zip(n, s))


def pre_diag_tp(cls, nf):
    """opfilt_tp.pre_op_diag (:87-104)"""
    lmax = len(nf.b) - 1
    fl = slinv_tp(cls, lmax)
    for i, f in enumerate(nf.ftebl()):
        fl[:, i, i] += f
    return np.linalg.pinv(fl)


def dot_tp(a, b):
    """opfilt_tp.dot_op (:46-58): all multipoles, T + E + B"""
    return sum(dot_tt(x, y) for x, y in zip(a, b))
