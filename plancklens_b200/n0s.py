"""Curved-sky reconstruction noise levels N0 of the quadratic estimators (reference: plancklens/n0s.py).

`get_N0` assembles, for an idealised experiment (Gaussian beam, white or scale-dependent noise), the filter spectra,
the spectra of the filtered maps, the unnormalised QE noise (`nhl.get_nhl`) and the responses (`qresp.get_response`);
all the Wigner small-d transforms underneath run on the GPU (`plk_wignerpos_dev` / `plk_wignercoeff_dev`).

Difference to the reference at the commit surveyed: its `get_N0` passes an undefined name `cls_glen` to the
separately-filtered responses (n0s.py:190) and raises `NameError` unless a module global of that name exists; here the
documented `cls_len` is used.  The golden vectors (tests/golden/make_golden_n0s.py) run the unmodified reference with
that global set to `cls_len`.
"""
import os
from copy import deepcopy

import numpy as np

from . import hp, nhl, qresp, utils

_CLS_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'cls')
_AMIN = np.pi / 180. / 60.


def _fiducial():
    return utils.camb_clfile(os.path.join(_CLS_PATH, 'FFP10_wdipole_lensedCls.dat'))


def _per_field(x, floor=None):
    """int or {'t':, 'e':, 'b':} -> dict"""
    if isinstance(x, dict):
        return x
    return {s: (x if floor is None else max(x, floor)) for s in 'teb'}


def _nlev_eb(nlev_p):
    """E and B noise levels from a scalar, an array, [array] or [array_E, array_B] (reference: n0s.py:73-97)."""
    if isinstance(nlev_p, (list, np.ndarray)):
        nlev_p = np.asarray(nlev_p)
        if nlev_p.shape[0] == 1:
            return nlev_p[0], nlev_p[0]
        if nlev_p.shape[0] == 2:
            return nlev_p[0], nlev_p[1]
        return nlev_p, nlev_p
    return nlev_p, nlev_p


def _band(cls, lmins, lmaxs, lmin_floor=0):
    """zeroes every spectrum outside the multipole range common to its two fields, in place"""
    for k, cl in cls.items():
        cl[min(lmaxs[k[0]], lmaxs[k[1]]) + 1:] = 0.
        cl[:max(lmin_floor, lmins[k[0]], lmins[k[1]])] = 0.


def get_N0(beam_fwhm=1.4, nlev_t=5., nlev_p=None, lmax_CMB=3000, lmin_CMB=100, lmax_out=None, cls_filt=None,
           cls_len=None, cls_weight=None, cls_sky=None, joint_TP=True, ksource='p', wfleg_Tcut=None):
    r"""N0 noise levels of the T-only, P-only and (G)MV estimators (reference: n0s.py:30-203).

        Args:
            beam_fwhm: Gaussian beam FWHM in arcmin
            nlev_t: temperature noise in uK-arcmin (scalar or array over multipoles)
            nlev_p: polarisation noise (defaults to sqrt(2) nlev_t; scalar, array, or (2, lmax + 1) for E and B)
            lmax_CMB, lmin_CMB: CMB multipole range (int or {'t', 'e', 'b'} dict)
            lmax_out: highest lensing multipole
            cls_filt, cls_len, cls_weight, cls_sky: filtering, response, QE-weight and data spectra (default FFP10 lensed)
            joint_TP: include the jointly filtered (GMV) estimator; otherwise the separately filtered MV
            ksource: anisotropy source
            wfleg_Tcut: high-l cut of T on the second (Wiener-filtered) leg

        Returns:
            (N0 gradient, N0 curl): dicts keyed by estimator
    """
    if nlev_p is None:
        nlev_p = nlev_t * np.sqrt(2)
    nlev_e, nlev_b = _nlev_eb(nlev_p)
    lmaxs = _per_field(lmax_CMB)
    lmins = _per_field(lmin_CMB, floor=1)
    lmax_ivf = int(np.max(list(lmaxs.values())))
    lmax_qlm = lmax_out or lmax_ivf
    cls_len = cls_len or _fiducial()
    cls_weight = cls_weight or _fiducial()
    cls_sky = cls_sky or _fiducial()
    cls_filt = cls_filt or _fiducial()

    qe_keys = [ksource + 'tt', ksource + '_p']
    if not joint_TP:
        qe_keys.append(ksource)

    transf = hp.gauss_beam(beam_fwhm * _AMIN, lmax=lmax_ivf)
    noise = {'tt': (nlev_t * _AMIN) ** 2 / transf ** 2, 'ee': (nlev_e * _AMIN) ** 2 / transf ** 2,
             'bb': (nlev_b * _AMIN) ** 2 / transf ** 2}

    def with_noise(src):
        out = {k: src[k][:lmax_ivf + 1] + noise[k] for k in ('tt', 'ee', 'bb')}
        out['te'] = np.copy(src['te'][:lmax_ivf + 1])
        _band(out, lmins, lmaxs)
        return out

    cls_dat, cls_filter = with_noise(cls_sky), with_noise(cls_filt)

    # separate filtering: 1 / (C + N) per field; joint filtering: the T E B matrix inverse
    fal_s = {k: utils.cli(cls_filter[k]) for k in ('tt', 'ee', 'bb')}
    fal_j = utils.cl_inverse(cls_filter)
    ivf = {}     # spectra of the filtered maps, fal . dat . fal^t, for the leg combinations aa, ab, ba, bb
    if wfleg_Tcut is not None and wfleg_Tcut < lmaxs['t']:
        fal_s_b = deepcopy(fal_s)
        fal_s_b['tt'][wfleg_Tcut + 1:] = 0.
        cut = deepcopy(cls_dat)
        for k in cut:
            if 't' in k:
                cut[k][wfleg_Tcut + 1:] = 0.
        fal_j_b = utils.cl_inverse(cut)
        for tag, (fa, fb) in (('s', (fal_s, fal_s_b)), ('j', (fal_j, fal_j_b))):
            ivf[tag] = {'aa': utils.cls_dot([fa, cls_dat, fa], ret_dict=True),
                        'ab': utils.cls_dot([fa, cls_dat, fb], ret_dict=True),
                        'ba': utils.cls_dot([fb, cls_dat, fa], ret_dict=True),
                        'bb': utils.cls_dot([fb, cls_dat, fb], ret_dict=True)}
    else:
        fal_s_b, fal_j_b = fal_s, fal_j
        for tag, fa in (('s', fal_s), ('j', fal_j)):
            aa = utils.cls_dot([fa, cls_dat, fa], ret_dict=True)
            ivf[tag] = {'aa': aa, 'ab': aa, 'ba': aa, 'bb': aa}
    seen = set()
    for cls in [fal_s, fal_j, fal_s_b, fal_j_b] + [c for tag in ('s', 'j') for c in ivf[tag].values()]:
        if id(cls) not in seen:
            seen.add(id(cls))
            _band(cls, lmins, {s: lmax_ivf for s in 'teb'}, lmin_floor=1)

    def n0(qe_key, cls_i, fal, fal_b):
        NG, NC, _, _ = nhl.get_nhl(qe_key, qe_key, cls_weight, cls_i['aa'], lmax_ivf, lmax_ivf, lmax_out=lmax_qlm,
                                   cls_ivfs_ab=cls_i['ab'], cls_ivfs_bb=cls_i['bb'], cls_ivfs_ba=cls_i['ba'])
        RG, RC, _, _ = qresp.get_response(qe_key, lmax_ivf, ksource, cls_weight, cls_len, fal, lmax_qlm=lmax_qlm,
                                          fal_leg2=fal_b)
        return utils.cli(RG ** 2) * NG, utils.cli(RC ** 2) * NC

    N0s, N0_curls = {}, {}
    for qe_key in qe_keys:
        N0s[qe_key], N0_curls[qe_key] = n0(qe_key, ivf['s'], fal_s, fal_s_b)
    if joint_TP:
        N0s[ksource], N0_curls[ksource] = n0(ksource, ivf['j'], fal_j, fal_j_b)
    return N0s, N0_curls


def cls2dls(cls):
    """Spectra dictionary -> CAMB array (lmax + 1, 4) of l(l+1)C_l/2pi in TT EE BB TE order, and [l(l+1)]^2 C^pp_l/2pi
    (reference: n0s.py:206-219)."""
    lmax = max(len(cl) for cl in cls.values()) - 1
    l = np.arange(lmax + 1, dtype=float)
    dls = np.zeros((lmax + 1, 4), dtype=float)
    for i, k in enumerate(('tt', 'ee', 'bb', 'te')):
        if k in cls:
            n = min(len(cls[k]), lmax + 1)
            dls[:n, i] = cls[k][:n] * (l * (l + 1) / (2. * np.pi))[:n]
    cldd = None
    if cls.get('pp', None) is not None:
        lp = np.arange(len(cls['pp']), dtype=float)
        cldd = cls['pp'] * (lp * (lp + 1)) ** 2 / (2. * np.pi)
    return dls, cldd


def dls2cls(dls):
    """Inverse of cls2dls for the CMB spectra (reference: n0s.py:222-230)."""
    assert dls.shape[1] == 4
    l = np.arange(dls.shape[0], dtype=float)
    fac = 2. * np.pi * utils.cli(l * (l + 1))
    return {k: dls[:, i] * fac for i, k in enumerate(('tt', 'ee', 'bb', 'te'))}


def get_N0_iter(*args, **kwargs):
    """Iterative-estimator N0 (reference: n0s.py:233-448).  It needs `camb.correlations.lensed_cls` for the partially
    delensed spectra at every iteration; camb is not part of this image, and no golden vector can be produced."""
    raise NotImplementedError("n0s.get_N0_iter needs the camb package (camb.correlations.lensed_cls)")
