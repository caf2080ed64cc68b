"""Generic quadratic-estimator evaluation from leg definitions (reference: plancklens/utils_qe.py).

`qe_eval` evaluates any estimator given as a list of `qe(leg_a, leg_b, cL)` terms: each leg is synthesised to a
spin-weighted map, legs are multiplied pixel by pixel, and the product is analysed back.  Here the legs, the
complex products and the analysis all stay on the GPU (one `alm_combine` per leg for the per-l weights, spin 0-3
synthesis, `plk_map_cmul_acc_dev`, spin-s analysis with the output scaling fused).
"""
import numpy as np
import torch

from . import hp, sht


class qeleg:
    """One leg: input spin, output spin and per-l weight (reference: utils_qe.py:5-33)."""

    def __init__(self, spin_in, spin_out, cl):
        self.spin_in = spin_in
        self.spin_ou = spin_out
        self.cl = cl

    def __eq__(self, leg):
        if self.spin_in != leg.spin_in or self.spin_ou != leg.spin_ou or self.get_lmax() != leg.get_lmax():
            return False
        return bool(np.all(self.cl == leg.cl))

    def __mul__(self, other):
        return qeleg(self.spin_in, self.spin_ou, self.cl * other)

    def __add__(self, other):
        assert self.spin_in == other.spin_in and self.spin_ou == other.spin_ou
        lmax = max(self.get_lmax(), other.get_lmax())
        cl = np.zeros(lmax + 1, dtype=np.result_type(self.cl, other.cl))
        cl[:len(self.cl)] += self.cl
        cl[:len(other.cl)] += other.cl
        return qeleg(self.spin_in, self.spin_ou, cl)

    def copy(self):
        return qeleg(self.spin_in, self.spin_ou, np.copy(self.cl))

    def get_lmax(self):
        return len(self.cl) - 1


def _dfl(fl):
    return torch.from_numpy(np.ascontiguousarray(fl, dtype=np.float64)).cuda()


class qeleg_multi:
    """Several (spin_in, cl) inputs sharing one output spin (reference: utils_qe.py:36-76)."""

    def __init__(self, spins_in, spin_out, cls):
        assert isinstance(spins_in, list) and isinstance(cls, list) and len(spins_in) == len(cls)
        self.spins_in = spins_in
        self.cls = cls
        self.spin_ou = spin_out

    def __iadd__(self, leg):
        assert leg.spin_ou == self.spin_ou, (leg.spin_ou, self.spin_ou)
        self.spins_in.append(leg.spin_in)
        self.cls.append(np.copy(leg.cl))
        return self

    def get_lmax(self):
        return int(np.max([len(cl) for cl in self.cls])) - 1

    def _weights(self):
        """per-l factors of T, E (gradient part) and B (curl part) after all the sign conventions of
        utils_qe.py:59-72: G = -sum_{s_in = 0} cl T + sum_{|s_in| = 2} cl E, C = +-sum sgn(s_in) cl B."""
        lmax = self.get_lmax()
        ft, fe, fb = np.zeros(lmax + 1), np.zeros(lmax + 1), np.zeros(lmax + 1)
        s_out = -1.0 if self.spin_ou > 0 else 1.0
        for si, cl in zip(self.spins_in, self.cls):
            assert si in [0, -2, 2], str(si) + ' input spin not implemented'
            assert not np.iscomplexobj(cl) or not np.any(np.imag(cl)), 'complex leg weights (TB / EB) are not on the GPU path'
            cl = np.real(cl)
            if si == 0:
                ft[:len(cl)] -= cl
            else:
                fe[:len(cl)] += cl
                fb[:len(cl)] += s_out * (1.0 if si < 0 else -1.0) * cl
        return ft, fe, fb

    def dev_map(self, get_dalm, nside):
        """(Re, Im) device maps of the leg; get_dalm(field, lmax) returns a device alm truncated to lmax."""
        lmax = self.get_lmax()
        ft, fe, fb = self._weights()
        plan = sht.get_plan(nside, lmax)
        terms = []
        if np.any(ft):
            terms.append((get_dalm('t', lmax), _dfl(ft)))
        if np.any(fe):
            terms.append((get_dalm('e', lmax), _dfl(fe)))
        from .qest import _combine
        glm = _combine(lmax, terms) if terms else torch.zeros(sht.alm_size(lmax), dtype=torch.complex128, device='cuda')
        clm = sht.almxfl(get_dalm('b', lmax), _dfl(fb)) if np.any(fb) else None
        s = abs(self.spin_ou)
        if s == 0:
            # uspin.alm2map_spin(spin 0) = (alm2map(-glm), 0)  (utils_spin.py:27)
            red = plan.alm2map(glm, fl=_dfl(-np.ones(lmax + 1)))
            return red, None
        red, imd = plan.alm2map_spin(glm, clm, s)
        if self.spin_ou < 0:
            neg = red if (self.spin_ou % 2 == 1) else imd
            sht.check(sht._lib.load().plk_alm_lincomb_dev(neg.numel() // 2, -1.0, sht._ptr(neg), 0.0, None, sht._ptr(neg), sht._stream()))
        return red, imd

    def __call__(self, get_alm, nside):
        """Spin-weighted real-space map of the leg as a complex numpy array (reference: utils_qe.py:50-73)."""
        red, imd = self.dev_map(_dalm_getter(get_alm), nside)
        r = red.cpu().numpy()
        return r + 1j * (imd.cpu().numpy() if imd is not None else 0.)


def _dalm_getter(get_alm):
    cache = {}

    def get(field, lmax):
        if field not in cache:
            cache[field] = sht.dev_alm(get_alm(field))
        a = cache[field]
        return a if sht.alm_lmax(a.numel()) == lmax else sht.alm_copy(a, lmax)
    return get


class qe:
    def __init__(self, leg_a, leg_b, cL):
        assert leg_a.spin_ou + leg_b.spin_ou >= 0
        self.leg_a = leg_a
        self.leg_b = leg_b
        self.cL = cL

    def get_lmax_a(self):
        return self.leg_a.get_lmax()

    def get_lmax_b(self):
        return self.leg_b.get_lmax()


def qe_eval(qe_list, nside, get_alm, lmax_qlm, verbose=True, get_alm2=None):
    """Gradient and curl alm of a QE given by its list of leg definitions (reference: utils_qe.py:92-132).

        Args:
            qe_list: list of qe instances
            nside: resolution of the real-space products
            get_alm: callable with 't', 'e', 'b' returning the inverse-variance filtered alms
            lmax_qlm: maximum multipole of the output
            get_alm2: alms of the second leg if different (the estimator is then symmetrised)
    """
    if get_alm2 is None:
        get_alm2 = get_alm
    symmetrize = get_alm2 is not get_alm
    qes = qe_compress(qe_list, verbose=verbose)
    qe_spin = qes[0][0].spin_ou + qes[0][1].spin_ou
    cL_out = qes[0][-1](np.arange(lmax_qlm + 1))
    assert qe_spin >= 0, qe_spin
    for q in qes[1:]:
        assert np.all(q[-1](np.arange(lmax_qlm + 1)) == cL_out)
        assert q[0].spin_ou + q[1].spin_ou == qe_spin
    npix = hp.nside2npix(nside)
    dr = torch.zeros(npix, dtype=torch.float64, device='cuda')
    di = torch.zeros(npix, dtype=torch.float64, device='cuda')
    g1, g2 = _dalm_getter(get_alm), _dalm_getter(get_alm2)
    for i, q in enumerate(qes):
        if verbose:
            print("QE %s out of %s :" % (i + 1, len(qes)))
            print("in-spins 1st leg and out-spin", q[0].spins_in, q[0].spin_ou)
            print("in-spins 2nd leg and out-spin", q[1].spins_in, q[1].spin_ou)
        pairs = [(g1, g2)] + ([(g2, g1)] if symmetrize else [])
        for ga, gb in pairs:
            ar, ai = q[0].dev_map(ga, nside)
            br, bi = q[1].dev_map(gb, nside)
            sht.map_cmul_acc(ar, ai, br, bi, dr, di)
    scale = cL_out * (0.5 if symmetrize else 1.0)
    plan = sht.get_plan(nside, lmax_qlm)
    if qe_spin > 0:
        fl = _dfl(scale)
        glm, clm = plan.map2alm_spin(dr, di, qe_spin, flg=fl, flc=fl)
        return glm.cpu().numpy(), clm.cpu().numpy()
    glm = plan.map2alm(dr, fl=_dfl(-scale))          # uspin.map2alm_spin(spin 0) = (-map2alm(re), 0)
    return glm.cpu().numpy(), 0.


def qe_proj(qe_list, a, b):
    """Restriction of a list of QEs to field a on the first leg and b on the second (reference: utils_qe.py:135-176)."""
    assert a in ['t', 'e', 'b'] and b in ['t', 'e', 'b']
    l_in = [0] if a == 't' else [-2, 2]
    r_in = [0] if b == 't' else [-2, 2]
    out = []
    for q in qe_list:
        si, ri = q.leg_a.spin_in, q.leg_b.spin_in
        if si not in l_in or ri not in r_in:
            continue
        la, lb = q.leg_a.copy(), q.leg_b.copy()
        sa = 1 if a == 'e' else -1
        sb = 1 if b == 'e' else -1
        if si == 0 and ri == 0:
            out.append(qe(la, lb, q.cL))
        elif si == 0:
            out.append(qe(la, lb * 0.5, q.cL))
            lb.spin_in *= -1
            out.append(qe(la, lb * 0.5 * sb, q.cL))
        elif ri == 0:
            out.append(qe(la * 0.5, lb, q.cL))
            la.spin_in *= -1
            out.append(qe(la * 0.5 * sa, lb, q.cL))
        else:
            out.append(qe(la * 0.5, lb * 0.5, q.cL))
            lb.spin_in *= -1
            out.append(qe(la * 0.5, lb * 0.5 * sb, q.cL))
            la.spin_in *= -1
            out.append(qe(la * 0.5 * sa, lb * 0.5 * sb, q.cL))
            lb.spin_in *= -1
            out.append(qe(la * 0.5 * sa, lb * 0.5, q.cL))
    return qe_simplify(out)


def qe_simplify(qe_list, _swap=False, verbose=False):
    """Co-adds terms that share a leg and differ only by the weight of the other (reference: utils_qe.py:179-205)."""
    skip = []
    ret = []
    qes = [qe(q.leg_b.copy(), q.leg_a.copy(), q.cL) for q in qe_list] if _swap else qe_list
    for i, q1 in enumerate(qes):
        if i in skip:
            continue
        leg_a, leg_b = q1.leg_a.copy(), q1.leg_b.copy()
        for j, q2 in enumerate(qes[i + 1:]):
            if q2.leg_a == leg_a and q2.leg_b.spin_in == q1.leg_b.spin_in and q2.leg_b.spin_ou == q1.leg_b.spin_ou:
                Ls = np.arange(max(q1.leg_b.get_lmax(), q2.leg_b.get_lmax()) + 1)
                if np.all(q1.cL(Ls) == q2.cL(Ls)):
                    leg_b = leg_b + q2.leg_b
                    skip.append(j + i + 1)
        if np.any(leg_a.cl) and np.any(leg_b.cl):
            ret.append(qe(leg_a, leg_b, q1.cL))
    if verbose and len(skip) > 0:
        print("%s terms down from %s" % (len(ret), len(qes)))
    if not _swap:
        return qe_simplify(ret, _swap=True, verbose=verbose)
    return [qe(q.leg_b.copy(), q.leg_a.copy(), q.cL) for q in ret]


def qe_compress(qes, verbose=True):
    """Merges terms with identical first leg so that fewer transforms are needed (reference: utils_qe.py:208-226)."""
    skip = []
    out = []
    for i, qi in enumerate(qes):
        if i in skip:
            continue
        lega_m = qeleg_multi([qi.leg_a.spin_in], qi.leg_a.spin_ou, [qi.leg_a.cl])
        legb_m = qeleg_multi([qi.leg_b.spin_in], qi.leg_b.spin_ou, [qi.leg_b.cl])
        for j, qj in enumerate(qes[i + 1:]):
            if qj.leg_a == qi.leg_a and legb_m.spin_ou == qj.leg_b.spin_ou:
                legb_m += qj.leg_b
                skip.append(i + 1 + j)
        out.append((lega_m, legb_m, qi.cL))
    if len(skip) > 0 and verbose:
        print("%s alm2map_spin transforms now required, down from %s" % (2 * (len(qes) - len(skip)), 2 * len(qes)))
    return out
