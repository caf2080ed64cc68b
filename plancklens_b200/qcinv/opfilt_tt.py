r"""Temperature-only Wiener / inverse-variance filter operators on the GPU (reference: plancklens/qcinv/opfilt_tt.py).

 :math:`S^{-1} (S^{-1} + Y^t N^{-1} Y)^{-1} Y^t N^{-1}`

Same plug-in surface as the reference module -- `calc_prep`, `apply_fini`, `dot_op`, `fwd_op`, `pre_op_diag`,
`pre_op_dense`, `alm_filter_ninv` -- so `multigrid.multigrid_chain(opfilt_tt, ...)` works unchanged.  Vectors are
`util_alm.dalm` (GPU resident); one `fwd_op` is: b_l scaling fused into the synthesis, N^{-1} and the
monopole/dipole projection as two per-pixel kernels, analysis with the b_l npix/4pi scaling fused, and a
two-term per-l combination for the C_l^{-1} term.  No host round trip inside the operator.
"""
import hashlib

import numpy as np
import torch

from .. import hp, sht
from ..utils import clhash
from . import dense, template_removal, util
from .util_alm import dalm


def _cli(cl):
    cl = np.asarray(cl, dtype=float)
    ret = np.zeros_like(cl)
    nz = cl != 0.
    ret[nz] = 1. / cl[nz]
    return ret


def _dev(x):
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).cuda()


def _as_dalm(alm):
    return alm if isinstance(alm, dalm) else dalm.from_numpy(alm)


def calc_prep(m, s_cls, n_inv_filt):
    """b = B^t N^{-1} d  (reference: opfilt_tt.py:30-36)."""
    tmap = sht.dev_map(m).clone() if isinstance(m, torch.Tensor) else sht.dev_map(m)
    n_inv_filt.apply_map(tmap)
    lmax = len(n_inv_filt.b_transf) - 1
    plan = sht.get_plan(n_inv_filt.nside, lmax)
    return dalm(plan.map2alm(tmap, fl=n_inv_filt.fl_out(lmax)), lmax)


def apply_fini(alm, s_cls, n_inv_filt):
    """Wiener-filtered solution -> inverse-variance filtered alm, in place (reference: opfilt_tt.py:39-41)."""
    fl = _dev(_cli(s_cls['tt']))
    if isinstance(alm, dalm):
        alm.almxfl(fl, inplace=True)
    else:
        alm[:] = hp.almxfl(alm, _cli(s_cls['tt']))


class dot_op:
    """sum_l (2l+1) C_l^{ab}  (reference: opfilt_tt.py:43-51)."""

    def __init__(self):
        pass

    def __call__(self, alm1, alm2):
        assert alm1.lmax == alm2.lmax
        return float(sht.alm_dot(alm1.t, alm2.t, lmin=0).item())

    def dev(self, alm1, alm2):
        """same number as a 1-element device tensor: no host synchronisation (fixed-iteration multigrid stages)"""
        assert alm1.lmax == alm2.lmax
        return sht.alm_dot(alm1.t, alm2.t, lmin=0)

    def fused(self, alm1, alm2, num=None, den=None, scale=1.0):
        """[s, r, -r] on the device in ONE kernel: s the dot product, r = scale * num / s or scale * s / den
        (the step lengths of cd_solve.py:69-71, :95-99)"""
        assert alm1.lmax == alm2.lmax
        return sht.alm_dot_fused([alm1.t], [alm2.t], lmin=0, num=num, den=den, scale=scale)


class fwd_op:
    """A x = C_l^{-1} x + B^t N^{-1} B x  (reference: opfilt_tt.py:54-73)."""

    def __init__(self, s_cls, n_inv_filt):
        self.cltt_inv = _cli(s_cls['tt'])
        self.n_inv_filt = n_inv_filt
        self._cltt_inv_d = _dev(self.cltt_inv)
        self._cli_d = {}

    def hashdict(self):
        return {'cltt_inv': clhash(self.cltt_inv), 'n_inv_filt': self.n_inv_filt.hashdict()}

    def __call__(self, talm):
        return self.calc(talm)

    def calc(self, talm):
        if talm.is_zero():   # do nothing if zero (reference: opfilt_tt.py:68)
            return talm
        # apply_alm (opfilt_tt.py:183-190) with the C_l^-1 x term folded into the output pass of the analysis
        nf = self.n_inv_filt
        lmax = talm.lmax
        plan = sht.get_plan(nf.nside, lmax)
        tmap = plan.alm2map(talm.t, fl=nf.fl_in(lmax))
        nf.apply_map(tmap)
        if lmax not in self._cli_d:
            self._cli_d[lmax] = sht.dev_fl(self.cltt_inv, lmax)
        return dalm(plan.map2alm_add(tmap, nf.fl_out(lmax), talm.t, self._cli_d[lmax]), lmax)


def _combine2(a, fla, b, flb):
    """fla[l] a + flb[l] b on device alms (fl None = 1) via plk_alm_combine_dev."""
    import ctypes
    lib = sht._lib.load()
    lmax = sht.alm_lmax(a.numel())
    ones = _ones(lmax)
    fla = ones if fla is None else fla
    flb = ones if flb is None else flb
    out = torch.empty_like(a)
    ins = (ctypes.c_void_p * 2)(a.data_ptr(), b.data_ptr())
    fls = (ctypes.c_void_p * 2)(fla.data_ptr(), flb.data_ptr())
    nfl = (ctypes.c_int * 2)(int(fla.numel()), int(flb.numel()))
    sht.check(lib.plk_alm_combine_dev(lmax, 2, ins, fls, nfl, sht._ptr(out), sht._stream()))
    return out


_ONES = {}


def _ones(lmax):
    if lmax not in _ONES:
        _ONES[lmax] = torch.ones(lmax + 1, dtype=torch.float64, device='cuda')
    return _ONES[lmax]


class pre_op_diag:
    """Harmonic-space diagonal preconditioner (reference: opfilt_tt.py:76-93)."""

    def __init__(self, s_cls, n_inv_filt):
        cltt = s_cls['tt']
        assert len(cltt) >= len(n_inv_filt.b_transf)
        n_inv_cl = np.sum(n_inv_filt.n_inv) / (4.0 * np.pi)
        lmax = len(n_inv_filt.b_transf) - 1
        assert lmax <= (len(cltt) - 1)
        filt = _cli(cltt[:lmax + 1])
        filt += n_inv_cl * n_inv_filt.b_transf[:lmax + 1] ** 2
        self.filt = _cli(filt)
        self._filt_d = _dev(self.filt)

    def __call__(self, talm):
        return self.calc(talm)

    def calc(self, talm):
        return talm.almxfl(self._filt_d)


def pre_op_dense(lmax, fwd_op, cache_fname=None):
    return dense.pre_op_dense_tt(lmax, fwd_op, cache_fname=cache_fname)


class alm_filter_ninv(object):
    """Pixel-space inverse noise with monopole / dipole marginalisation (reference: opfilt_tt.py:99-205)."""

    def __init__(self, n_inv, b_transf, marge_monopole=False, marge_dipole=False, marge_uptolmin=-1, marge_maps=(),
                 nlev_ftl=None):
        if isinstance(n_inv, list):
            n_inv_prod = util.load_map(n_inv[0])
            for n in n_inv[1:]:
                n_inv_prod = n_inv_prod * util.load_map(n)
            n_inv = n_inv_prod
        else:
            n_inv = util.load_map(n_inv)
        n_inv = np.asarray(n_inv, dtype=float)
        if marge_uptolmin >= 0:
            raise NotImplementedError("marge_uptolmin is not on the GPU path; monopole, dipole and template-map "
                                      "marginalisation are (SURVEY.md section 8a)")
        nz = n_inv != 0.0
        print("opfilt_tt: inverse noise map std dev / av = %.3e" % (np.std(n_inv[nz]) / np.average(n_inv[nz])))

        self.n_inv = n_inv
        self.b_transf = np.asarray(b_transf, dtype=float)
        self.npix = len(n_inv)
        self.nside = hp.npix2nside(self.npix)
        self.marge_monopole = marge_monopole
        self.marge_dipole = marge_dipole
        self.marge_uptolmin = marge_uptolmin
        self.templates = []
        self.templates_hash = []
        for tmap in [np.asarray(util.load_map(m), dtype=float) for m in marge_maps]:   # reference order: maps first
            assert len(n_inv) == len(tmap)
            self.templates.append(template_removal.template_map(tmap))
            self.templates_hash.append(hashlib.sha1(np.ascontiguousarray(tmap).view(np.uint8)).hexdigest())
        if marge_monopole:
            self.templates.append(template_removal.template_monopole())
        if marge_dipole:
            self.templates.append(template_removal.template_dipole())

        self._ninv_d = _dev(n_inv)
        self._fl_cache = {}
        self._plan0 = sht.get_plan(self.nside, max(len(self.b_transf) - 1, 1))
        # Internal mode order: (1, x, y, z) analytic modes first (slots of unused ones stay empty), then the template
        # maps; `Pt_Nn1_P_inv` is kept in the reference's order (maps, monopole, dipole).
        self._tmaps = [t for t in self.templates if isinstance(t, template_removal.template_map)]
        nt = 4 + len(self._tmaps)
        self._sums = torch.zeros(nt, dtype=torch.float64, device='cuda')
        if len(self.templates) != 0:
            amodes = [i for t in self.templates for i in t.modes]
            act = amodes + [4 + i for i in range(len(self._tmaps))]
            # P^t N^{-1} P: row a = sum_p n_inv mode_a {1, x, y, z, tmap_0, ...}
            full = np.zeros((nt, nt))
            minus_eye = -torch.eye(4, dtype=torch.float64, device='cuda')
            self._ntm = []                       # n_inv * template map, what apply_map subtracts
            for a in act:
                if a < 4:
                    ea = torch.zeros(4, dtype=torch.float64, device='cuda')
                    ea[a] = 1.0
                    tmp = torch.zeros(self.npix, dtype=torch.float64, device='cuda')
                    self._plan0.modes_sub(tmp, self._ninv_d, ea, minus_eye.reshape(-1))   # tmp = n_inv * mode_a
                else:
                    tmp = sht.map_mul(self._tmaps[a - 4].map.clone(), self._ninv_d)
                    self._ntm.append(tmp)
                full[a, :4] = self._plan0.modes_dot(tmp).cpu().numpy()
                for j, t in enumerate(self._tmaps):
                    full[a, 4 + j] = float(sht.map_dot(t.map, tmp).item())
            sub = full[np.ix_(act, act)]
            sub = 0.5 * (sub + sub.T)
            eigv, eigw = np.linalg.eigh(sub)
            inv = np.dot(np.dot(eigw, np.diag(1.0 / eigv)), np.transpose(eigw))
            ref_order = [act.index(4 + i) for i in range(len(self._tmaps))] + [act.index(a) for a in amodes]
            self.Pt_Nn1_P_inv = inv[np.ix_(ref_order, ref_order)]
            pinv = np.zeros((nt, nt))
            pinv[np.ix_(act, act)] = inv
            if len(self._tmaps) == 0:
                self._pinv_d = _dev(pinv.reshape(-1))
            else:
                self._pinv_d = _dev(pinv.reshape(-1))
                self._mpinv_d = _dev(-pinv.reshape(-1))
                self._eye4_d = torch.eye(4, dtype=torch.float64, device='cuda').reshape(-1).contiguous()
                self._coef = torch.zeros(nt, dtype=torch.float64, device='cuda')
                self._mcoef = torch.zeros(nt, dtype=torch.float64, device='cuda')

        if nlev_ftl is None:
            nlev_ftl = 10800. / np.sqrt(np.sum(self.n_inv) / (4.0 * np.pi)) / np.pi
        self.nlev_ftl = nlev_ftl
        print("ninv_ftl: using %.2f uK-amin noise Cl" % self.nlev_ftl)

    def hashdict(self):
        return {'n_inv': clhash(self.n_inv), 'b_transf': clhash(self.b_transf),
                'marge_monopole': self.marge_monopole, 'marge_dipole': self.marge_dipole,
                'templates_hash': self.templates_hash, 'marge_uptolmin': self.marge_uptolmin}

    def get_ftl(self):
        return self.b_transf ** 2 / (self.nlev_ftl / 60. / 180. * np.pi) ** 2

    def degrade(self, nside):
        """Same filter on a coarser grid: n_inv summed over children (reference: opfilt_tt.py:172-181)."""
        if nside == self.nside:
            return self
        print("DEGRADING WITH NO MARGE MAPS")
        return alm_filter_ninv(sht.ud_grade_sum(self._ninv_d, nside).cpu().numpy(), self.b_transf,
                               marge_monopole=self.marge_monopole, marge_dipole=self.marge_dipole,
                               marge_uptolmin=self.marge_uptolmin, marge_maps=[])

    # per-l factors on the device, padded to the lmax of the vector they multiply
    def fl_in(self, lmax):
        k = ('in', lmax)
        if k not in self._fl_cache:
            self._fl_cache[k] = sht.dev_fl(self.b_transf, lmax)
        return self._fl_cache[k]

    def fl_out(self, lmax):
        k = ('out', lmax)
        if k not in self._fl_cache:
            self._fl_cache[k] = sht.dev_fl(self.b_transf * (self.npix / (4. * np.pi)), lmax)
        return self._fl_cache[k]

    def apply_alm(self, alm):
        """alm <- B^t N^{-1} B alm, in place (reference: opfilt_tt.py:183-190)."""
        host = not isinstance(alm, dalm)
        v = _as_dalm(alm)
        plan = sht.get_plan(self.nside, v.lmax)
        tmap = plan.alm2map(v.t, fl=self.fl_in(v.lmax))
        self.apply_map(tmap)
        plan.map2alm(tmap, fl=self.fl_out(v.lmax), out=v.t)
        v.zero = False
        if host:
            alm[:] = v.numpy()

    def apply_map(self, tmap):
        """tmap <- N^{-1} tmap with the templates projected out, in place (reference: opfilt_tt.py:193-205)."""
        host = not isinstance(tmap, torch.Tensor)
        t = sht.dev_map(tmap) if host else tmap
        plan = sht.get_plan(self.nside, self._plan0.lmax)
        if len(self.templates) != 0 and len(self._tmaps) == 0:
            plan.modes_dot(t, w=self._ninv_d, out=self._sums)          # t *= n_inv ; sums = P^t t
            plan.modes_sub(t, self._ninv_d, self._sums, self._pinv_d)    # t -= n_inv P (P^t N^-1 P)^-1 sums
        elif len(self.templates) != 0:
            # with template maps (reference: opfilt_tt.py:193-205 with template_removal.template_map): all on the
            # device, no host synchronisation
            lib = sht._lib.load()
            nt = self._sums.numel()
            plan.modes_dot(t, w=self._ninv_d, out=self._sums[:4])      # t *= n_inv ; analytic sums
            for j, tm in enumerate(self._tmaps):
                sht.map_dot(tm.map, t, out=self._sums[4 + j:5 + j])
            sht.check(lib.plk_dense_matvec_dev(nt, sht._ptr(self._pinv_d), sht._ptr(self._sums), sht._ptr(self._coef), sht._stream()))
            sht.check(lib.plk_dense_matvec_dev(nt, sht._ptr(self._mpinv_d), sht._ptr(self._sums), sht._ptr(self._mcoef), sht._stream()))
            plan.modes_sub(t, self._ninv_d, self._coef[:4], self._eye4_d)
            for j, ntm in enumerate(self._ntm):
                sht.map_axpy_dev(t, ntm, self._mcoef[4 + j:5 + j])
        else:
            sht.map_mul(t, self._ninv_d)
        if host:
            tmap[:] = t.cpu().numpy()
