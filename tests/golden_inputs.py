"""Seeded small inputs shared by tests/golden/make_golden.py (reference side) and the parity tests (CUDA side)."""
import numpy as np


def alm_size(lmax):
    return (lmax + 1) * (lmax + 2) // 2


def alm_ls(lmax):
    return np.concatenate([np.arange(m, lmax + 1) for m in range(lmax + 1)])


def rand_alm(rng, lmax, lmin=0):
    n = alm_size(lmax)
    a = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    a[:lmax + 1] = a[:lmax + 1].real
    a[alm_ls(lmax) < lmin] = 0
    return a


def toy_cls(lmax):
    l = np.arange(lmax + 1, dtype=float)
    lp = np.maximum(l, 2.0)
    tt = 1e3 / (lp * (lp + 1)) * np.exp(-(lp / 60.0) ** 2) + 1e-4
    ee = 30.0 / (lp * (lp + 1)) * np.exp(-(lp / 70.0) ** 2) + 1e-5
    bb = 0.1 * ee
    te = 0.4 * np.sqrt(tt * ee) * np.cos(lp / 9.0)
    for c in (tt, ee, bb, te):
        c[:2] = 0.0
    return {'tt': tt, 'ee': ee, 'bb': bb, 'te': te}


def pix_z(nside):
    """z = cos(theta) of every RING pixel (ring geometry only)."""
    N = nside
    z = np.empty(12 * N * N)
    p = 0
    for i in range(1, 4 * N):
        if i < N:
            n, zz = 4 * i, 1 - i * i / (3.0 * N * N)
        elif i <= 3 * N:
            n, zz = 4 * N, (2 * N - i) * 2.0 / (3.0 * N)
        else:
            ip = 4 * N - i
            n, zz = 4 * ip, -(1 - ip * ip / (3.0 * N * N))
        z[p:p + n] = zz
        p += n
    return z


def cg_case(nside=32, lmax=64, seed=1234):
    rng = np.random.default_rng(seed)
    npix = 12 * nside ** 2
    z = pix_z(nside)
    mask = (np.abs(z) > np.sin(np.deg2rad(15.0))).astype(float)
    holes = rng.choice(npix, 40, replace=False)
    mask[holes] = 0.0
    cls = toy_cls(lmax)
    transf = np.exp(-0.5 * np.arange(lmax + 1) * (np.arange(lmax + 1) + 1.0) * (0.03 ** 2))
    pixarea = 4 * np.pi / npix
    # pixel noise chosen so that S/N crosses one near l ~ 40 (a few tens of CG iterations, as in production)
    ninv_t = mask * (1.0 + 0.5 * z ** 2) / 300.0 * (pixarea / (4 * np.pi / (12 * 32 ** 2)))
    ninv_p = mask * (1.0 + 0.3 * z ** 2) / 10.0
    return {'nside': nside, 'lmax': lmax, 'cls': cls, 'transf': transf,
            'ninv_t': [ninv_t], 'ninv_p1': [[ninv_p]],
            'ninv_p3': [[ninv_p], [0.2 * ninv_p * z], [ninv_p * (1.0 + 0.1 * z)]],
            'tmap': rng.standard_normal(npix) * 3.0, 'qmap': rng.standard_normal(npix), 'umap': rng.standard_normal(npix),
            'x_t': rand_alm(rng, lmax), 'x_e': rand_alm(rng, lmax, 2), 'x_b': rand_alm(rng, lmax, 2)}


def chain_descr_t(cd_solve):
    """Two-level version of the reference's default T chain (filt_cinv.py:113-116), sized for nside 32."""
    return [[1, ["split(dense, 8, diag_cl)"], 32, 16, 3, 0.0, cd_solve.tr_cg, cd_solve.cache_mem()],
            [0, ["split(stage(1), 32, diag_cl)"], 64, 32, np.inf, 1.0e-6, cd_solve.tr_cg, cd_solve.cache_mem()]]


def chain_descr_p(cd_solve):
    return [[1, ["split(dense, 8, diag_cl)"], 32, 16, 3, 0.0, cd_solve.tr_cg, cd_solve.cache_mem()],
            [0, ["split(stage(1), 32, diag_cl)"], 64, 32, np.inf, 1.0e-6, cd_solve.tr_cg, cd_solve.cache_mem()]]


def qe_case(nside=32, lmax=48, lmax_qlm=64, seed=4321):
    rng = np.random.default_rng(seed)
    cls = toy_cls(lmax)
    q = {'nside': nside, 'lmax': lmax, 'lmax_qlm': lmax_qlm, 'cls': cls}
    for tag in ('1', '2'):
        q['tlm' + tag] = rand_alm(rng, lmax, 2) * 1e-2
        q['elm' + tag] = rand_alm(rng, lmax, 2) * 1e-1
        q['blm' + tag] = rand_alm(rng, lmax, 2) * 1e-1
    return q


def template_case(nside=32, seed=99):
    """Pixel-space templates for the marginalisation tests: two T maps, one Q map, two U maps (seeded)."""
    rng = np.random.default_rng(seed)
    npix = 12 * nside ** 2
    z = pix_z(nside)
    return {'tmaps': [z ** 2 + 0.1 * rng.standard_normal(npix), np.cos(7 * z) + 0.3 * rng.standard_normal(npix)],
            'qmaps': [1.0 + 0.5 * z + 0.2 * rng.standard_normal(npix)],
            'umaps': [z ** 3 + 0.2 * rng.standard_normal(npix), rng.standard_normal(npix)]}


def chain_descr_tp(cd_solve):
    """Two-level version of the reference's default joint T+P chain (filt_cinv.py:398-405), sized for nside 32."""
    return [[1, ["split(dense, 8, diag_cl)"], 32, 16, 3, 0.0, cd_solve.tr_cg, cd_solve.cache_mem()],
            [0, ["split(stage(1), 32, diag_cl)"], 64, 32, np.inf, 1.0e-6, cd_solve.tr_cg, cd_solve.cache_mem()]]


class idx_ivfs:
    """In-memory filtering library whose simulations differ by index (deterministic mixes of the two seeded sets of
    qe_case); `hp` is the healpy-shaped module providing almxfl (the shim on the reference side, plancklens_b200.hp
    on the product side)."""
    lib_dir = None

    def __init__(self, q, hp, shift=0):
        self.q, self.hp, self.shift = q, hp, shift
        self.cl = q['cls']

    def hashdict(self):
        return {'idx_ivfs': 1, 'shift': self.shift}

    def get_tal(self, a):
        l = np.arange(self.q['lmax'] + 1, dtype=float)
        return np.exp(0.5 * l * (l + 1) * 0.01 ** 2)          # inverse of a Gaussian transfer function

    def get_fmask(self):
        return np.ones(12 * self.q['nside'] ** 2)

    def _mix(self, a, idx):
        i = (idx if idx >= 0 else 7) + self.shift
        return self.q[a + 'lm1'] * (1.0 + 0.3 * i) + self.q[a + 'lm2'] * (0.2 * i - 0.1)

    def get_sim_tlm(self, idx): return self._mix('t', idx)
    def get_sim_elm(self, idx): return self._mix('e', idx)
    def get_sim_blm(self, idx): return self._mix('b', idx)
    def get_sim_tmliklm(self, idx): return self.hp.almxfl(self.get_sim_tlm(idx), self.q['cls']['tt'])
    def get_sim_emliklm(self, idx): return self.hp.almxfl(self.get_sim_elm(idx), self.q['cls']['ee'])
    def get_sim_bmliklm(self, idx): return self.hp.almxfl(self.get_sim_blm(idx), self.q['cls']['bb'])


def resp_case(lmax=60, lmax_qlm=70):
    """Spectra, filters and filtered-map spectra for the response / N0 tests (toy spectra, 30' beam, lmin 10):
    separately filtered, jointly filtered, and a joint set with TB / EB correlations (complex spin matrices)."""
    cls = toy_cls(lmax)
    l = np.arange(lmax + 1, dtype=float)
    transf = np.exp(-0.5 * l * (l + 1) * (0.01 ** 2))
    nlt, nlp = 2e-3 / transf ** 2, 4e-3 / transf ** 2

    def cli(c):
        r = np.zeros_like(c)
        r[c != 0] = 1. / c[c != 0]
        return r
    fal_sep = {'tt': cli(cls['tt'] + nlt), 'ee': cli(cls['ee'] + nlp), 'bb': cli(cls['bb'] + nlp)}
    cls_ivfs_sep = {'tt': fal_sep['tt'].copy(), 'ee': fal_sep['ee'].copy(), 'bb': fal_sep['bb'].copy(),
                    'te': cls['te'] * fal_sep['tt'] * fal_sep['ee']}
    cls_dat = {'tt': cls['tt'] + nlt, 'ee': cls['ee'] + nlp, 'bb': cls['bb'] + nlp, 'te': cls['te'].copy()}

    def inv3(d):
        m = np.zeros((lmax + 1, 3, 3))
        for k, (i, j) in zip(['tt', 'ee', 'bb', 'te', 'tb', 'eb'], [[0, 0], [1, 1], [2, 2], [0, 1], [0, 2], [1, 2]]):
            if k in d:
                m[:, i, j] = m[:, j, i] = d[k]
        mi = np.linalg.pinv(m)
        return {k: mi[:, i, j].copy() for k, (i, j) in zip(['tt', 'ee', 'bb', 'te', 'tb', 'eb'],
                                                          [[0, 0], [1, 1], [2, 2], [0, 1], [0, 2], [1, 2]]) if np.any(mi[:, i, j])}
    fal_jt = inv3(cls_dat)
    dat_tb = dict(cls_dat)
    dat_tb['tb'] = 0.1 * np.sqrt(cls['tt'] * cls['bb'])
    dat_tb['eb'] = 0.05 * np.sqrt(cls['ee'] * cls['bb'])
    fal_tb = inv3(dat_tb)
    out = {'lmax': lmax, 'lmax_qlm': lmax_qlm, 'cls_weight': cls, 'cls_len': cls, 'cls_dat': cls_dat,
           'fal_sep': fal_sep, 'fal_jt': fal_jt, 'fal_tb': fal_tb,
           'cls_ivfs_sep': cls_ivfs_sep, 'cls_ivfs_jt': inv3(cls_dat), 'cls_ivfs_tb': inv3(dat_tb)}
    for d in (fal_sep, fal_jt, fal_tb, cls_ivfs_sep, out['cls_ivfs_jt'], out['cls_ivfs_tb']):
        for v in d.values():
            v[:10] = 0.
    return out


def kk_cls(lmax):
    """toy lensing-potential spectrum in the units of cg_case's noise: C_L^{kk} = (L(L+1)/2)^2 C_L^{pp} crosses the
    noise level of cg_case()['ninv_t'] near L ~ 45 (tens of CG iterations); zero below L = 2"""
    l = np.maximum(np.arange(lmax + 1, dtype=float), 1.0)
    clkk = 30.0 / (1.0 + (l / 10.0) ** 2) ** 1.5
    clpp = clkk / (0.5 * l * (l + 1)) ** 2
    clpp[:2] = 0.0
    return {'pp': clpp}


def n0s_cases(lmax=60, lmax_out=70):
    """Keyword arguments of n0s.get_N0: toy spectra (response spectra differing from the weights), 80' beam."""
    cls = toy_cls(lmax)
    cls_len = {k: v * (1.0 + 0.05 * np.cos(np.arange(lmax + 1) / 7.0)) for k, v in cls.items()}
    cls_sky = {k: v * 1.02 for k, v in cls.items()}
    base = dict(beam_fwhm=80., nlev_t=150., lmax_CMB=lmax, lmin_CMB=2, lmax_out=lmax_out, cls_filt=cls,
                cls_len=cls_len, cls_weight=cls, cls_sky=cls_sky)
    l = np.arange(lmax + 1, dtype=float)
    return {'gmv': dict(base, nlev_p=210., joint_TP=True),
            'sep': dict(base, joint_TP=False, lmax_CMB={'t': 50, 'e': lmax, 'b': lmax},
                        nlev_p=np.array([200. + l, 230. + 0.5 * l])),
            'tcut': dict(base, nlev_p=210., joint_TP=True, wfleg_Tcut=40, ksource='p'),
            'curl': dict(base, nlev_p=[210. * np.ones(lmax + 1)], joint_TP=True, ksource='x', lmin_CMB={'t': 5, 'e': 3, 'b': 3})}


QEST_EXTRA_KEYS = ['stt', 'ntt', 'ftt', 'f_p', 'f', 'a_p', 'pte', 'pet', 'xbt', 'peb', 'pbe', 'xee', 'ptb', 'p_eb', 'x_te',
                   'f_tp', 'ptt_bh_s', 's']


class toy_resplib:
    """deterministic stand-in for qresp.resp_lib_simple: enough for the bias-hardening combination of qest.library"""

    def __init__(self, lmax_qlm):
        self.L = np.arange(lmax_qlm + 1, dtype=float)

    def get_response(self, k, ksource):
        seed = sum(ord(c) for c in k + ksource)
        r = 1.0 + 0.01 * seed + 0.3 * np.cos(self.L / (3.0 + seed % 5))
        r[:2] = 0.0
        return r


def wrapper_case(q):
    """filters and spectra for the host-side wrapper goldens (make_golden_wrappers.py)"""
    lmax = q['lmax']
    l = np.arange(lmax + 1, dtype=float)
    transf = np.exp(-0.5 * l * (l + 1) * 0.02 ** 2)
    return {'lmax_cut': lmax - 6, 'transf': transf,
            'ftl': 1.0 / (1.0 + 0.01 * l), 'fel': 2.0 / (1.0 + 0.02 * l), 'fbl': 3.0 / (1.0 + 0.03 * l),
            'fm_t': 1.0 - 0.5 * np.exp(-l / 4.0), 'fm_e': 1.0 - 0.3 * np.exp(-l / 6.0), 'fm_b': np.where(l < 3, 0.0, 1.0),
            'nl_t': 1e-4 * (1 + 30.0 / (l + 1)), 'nl_p': 3e-4 * np.ones(lmax + 1)}


class alm_sims:
    """simulation library handing out harmonic coefficients (what library_fullsky_alms_sepTP and
    cmb_maps_harmonicspace consume), deterministic mixes of qe_case's alms"""

    def __init__(self, q):
        self.q, self.lmax = q, q['lmax']

    def hashdict(self):
        return {'alm_sims': 1}

    def _mix(self, a, idx):
        return self.q[a + 'lm1'] * (1.0 + 0.5 * idx) - self.q[a + 'lm2'] * 0.25 * idx

    def get_sim_tlm(self, idx): return self._mix('t', idx)
    def get_sim_elm(self, idx): return self._mix('e', idx)
    def get_sim_blm(self, idx): return self._mix('b', idx)
    def get_sim_tmap(self, idx): return self._mix('t', idx)
    def get_sim_pmap(self, idx): return self._mix('e', idx), self._mix('b', idx)


class map_sims:
    """tiny map-valued simulation library (12 pixels), value = scale * f(idx, pixel)"""

    def __init__(self, q, scale):
        self.scale = scale

    def hashdict(self):
        return {'map_sims': self.scale}

    def get_sim_tmap(self, idx):
        return self.scale * (np.arange(12.) + idx)

    def get_sim_pmap(self, idx):
        return self.scale * np.cos(np.arange(12.) + idx), self.scale * np.sin(np.arange(12.) * idx)


class fixed_phas:
    """harmonic phase library with three deterministic fields (stand-in for sims.phas.lib_phas)"""
    nfields = 3

    def __init__(self, q):
        self.q, self.lmax = q, q['lmax']

    def hashdict(self):
        return {'fixed_phas': 1}

    def get_sim(self, idx, idf=None):
        return [self.q['tlm2'], self.q['elm2'], self.q['blm2']][idf] * (1.0 + idx)


RESP2_CUSTOM = [('ptt', 'n'), ('ntt', 'n'), ('stt', 'ntt'), ('ftt', 'n')]
RESP2_DERIV = [('ptt', 30, 'tt', 'p'), ('p_p', 25, 'ee', 'p'), ('p', 40, 'te', 'p')]


def resp_transf(lmax):
    l = np.arange(lmax + 1, dtype=float)
    return np.exp(-0.5 * l * (l + 1) * (0.01 ** 2))


def apo_case(nside=16, lmax=40):
    """binary mask (galactic-like band + a few holes), a seeded map, and a smoothing scale of a few pixels"""
    rng = np.random.default_rng(99)
    z = pix_z(nside)
    mask = (np.abs(z) > 0.3).astype(float)
    mask[rng.choice(mask.size, 15, replace=False)] = 0.0
    return {'nside': nside, 'lmax': lmax, 'mask': mask, 'map': rng.standard_normal(mask.size), 'sigma_arcmin': 400.0}


N1_PAIRS = [('ptt', 'ptt'), ('pee', 'ptt'), ('ptt', 'pee'), ('p_p', 'p_p'), ('p', 'ptt'), ('ptt', 'p_tp'), ('p_eb', 'p_eb'),
            ('x_p', 'xtt'), ('stt', 'stt')]


def n1_case(lmax=80, lmaxphi=120, Lmax=60):
    cls = toy_cls(lmax)
    l = np.arange(lmax + 1, dtype=float)
    cut = (l >= 5).astype(float)
    return {'cltt': cls['tt'], 'clte': cls['te'], 'clee': cls['ee'], 'lmaxphi': lmaxphi, 'Lmax': Lmax,
            'clpp': 1e-7 / (1.0 + np.arange(lmaxphi + 1.0)) ** 2,
            'ftl': cut / (1.0 + 0.01 * l), 'fel': cut * 2.0 / (1.0 + 0.02 * l), 'fbl': cut * 3.0 / (1.0 + 0.03 * l),
            'ftlB': np.where(l >= 8, 1.0, 0.0) / (1.0 + 0.015 * l)}


def fake_n1l(L, cl_kind, kA, kB, k_ind, cltt, clte, clee, clttfid, cltefid, cleefid, ftlA, felA, fblA, ftlB, felB, fblB,
             lminA, lminB, dL, lps):
    """deterministic stand-in for the Fortran n1f.n1l: depends on every argument a wrong call order would change"""
    code = lambda k: sum((i + 1) * ord(ch) for i, ch in enumerate(k))
    return (1e-3 * (code(kA) + 0.37 * code(kB) + 0.11 * ord(k_ind)) * (1.0 + np.cos(L / 9.0)) / (L + 3.0)
            * (1.0 + np.sum(ftlA) + 2 * np.sum(felA) + 3 * np.sum(fblA) + 5 * np.sum(ftlB) + 7 * np.sum(felB) + 11 * np.sum(fblB))
            * (1.0 + lminA + 0.5 * lminB) * (1.0 + 0.01 * dL + 1e-3 * len(lps)) * (1.0 + 1e3 * cl_kind[min(int(L), len(cl_kind) - 1)]))


# ---------------------------------------------------------------------------------------------------------------------
# Config-3 filter libraries (filt_cinv.cinv_t / cinv_p / library_cinv_sepTP) at the smallest size their constructors
# accept (nside 512, lmax 1024: reference filt_cinv.py:77, :228) with the DEFAULT chains.
def camb_cls(path, lmax):
    """columns l, TT, EE, BB, TE in l(l+1)C_l/2pi (reference utils.camb_clfile, utils.py:308-333), restated so that both
    sides read the same numbers without importing each other's loader"""
    d = np.loadtxt(path).T
    ell = np.int_(d[0])
    w = ell * (ell + 1) / (2. * np.pi)
    cls = {}
    for k, col in (('tt', 1), ('ee', 2), ('bb', 3), ('te', 4)):
        c = np.zeros(lmax + 1)
        idc = ell <= lmax
        c[ell[idc]] = d[col][idc] / w[idc]
        cls[k] = c
    return cls


def cinv_case(alm2map, alm2map_spin, clpath, nside=512, lmax=1024, seed=512, nlev_t=35., nlev_p=55.):
    """Masked anisotropic-noise sky of SURVEY.md section 8d: |z| < sin 20 deg cut plus 400 seeded discs of ~15' radius,
    n_inv = mask (vamin / nlev)^2 (1 + 0.5 z^2), 5' beam, 35 / 55 uK-arcmin, data = Gaussian CMB (fiducial lensed
    spectra) * beam + white noise drawn in pixel space.  `alm2map`, `alm2map_spin` are the CPU oracle's on both the
    reference side (tests/golden/make_golden_cinv.py) and the test side, so both see bit-identical maps.
    The "deep" variant (nlev_t 3, nlev_p 4 uK-arcmin) is signal dominated to l ~ lmax: several tens of top-level
    iterations, past the `roundoff = 25` residual refresh of cd_solve.py:79-81."""
    rng = np.random.default_rng(seed)
    npix = 12 * nside ** 2
    cls = camb_cls(clpath, lmax)
    ell = np.arange(lmax + 1, dtype=float)
    sigma = (5. / 60. / 180. * np.pi) / np.sqrt(8. * np.log(2.))
    transf = np.exp(-0.5 * ell * (ell + 1) * sigma ** 2)
    z = pix_z(nside)
    mask = (np.abs(z) >= np.sin(np.deg2rad(20.))).astype(float)
    # discs: pixels within 15' of seeded centres, from ring geometry (phi of each pixel)
    phi = pix_phi(nside)
    sth = np.sqrt(1. - z ** 2)
    vec = np.stack([sth * np.cos(phi), sth * np.sin(phi), z])
    zc = rng.uniform(-1, 1, 400)
    pc = rng.uniform(0, 2 * np.pi, 400)
    sc = np.sqrt(1 - zc ** 2)
    cen = np.stack([sc * np.cos(pc), sc * np.sin(pc), zc])
    cosr = np.cos(np.deg2rad(15. / 60.))
    for j in range(400):
        band = np.abs(z - zc[j]) < 0.01           # cheap pre-cut
        idx = np.where(band)[0]
        hit = idx[(cen[:, j] @ vec[:, idx]) > cosr]
        mask[hit] = 0.
    vamin = np.sqrt(4. * np.pi / npix) * 180. * 60. / np.pi
    ninv_t = mask * (vamin / nlev_t) ** 2 * (1. + 0.5 * z ** 2)
    ninv_p = mask * (vamin / nlev_p) ** 2 * (1. + 0.5 * z ** 2)
    # CMB: independent T, E, B phases coloured by TT / EE / BB (TE correlation is irrelevant to separate filtering)
    ls = alm_ls(lmax)
    tlm = rand_alm(rng, lmax) * np.sqrt(0.5 * cls['tt'])[ls] * transf[ls]
    elm = rand_alm(rng, lmax, 2) * np.sqrt(0.5 * cls['ee'])[ls] * transf[ls]
    blm = rand_alm(rng, lmax, 2) * np.sqrt(0.5 * cls['bb'])[ls] * transf[ls]
    for a in (tlm, elm, blm):
        a[:lmax + 1] *= np.sqrt(2.)
    tmap = np.asarray(alm2map(tlm, nside)) + rng.standard_normal(npix) * (nlev_t / vamin)
    q, u = alm2map_spin([elm, blm], nside, 2, lmax)
    qmap = np.asarray(q) + rng.standard_normal(npix) * (nlev_p / vamin)
    umap = np.asarray(u) + rng.standard_normal(npix) * (nlev_p / vamin)
    return {'nside': nside, 'lmax': lmax, 'cls': cls, 'transf': transf, 'mask': mask,
            'ninv_t': [ninv_t], 'ninv_p': [[ninv_p]], 'tmap': tmap, 'qmap': qmap, 'umap': umap}


def pix_phi(nside):
    """phi of every RING pixel (ring geometry only)."""
    N = nside
    phi = np.empty(12 * N * N)
    p = 0
    for i in range(1, 4 * N):
        if i < N:
            n, sh = 4 * i, 1
        elif i <= 3 * N:
            n, sh = 4 * N, 1 if (i - N) % 2 == 0 else 0
        else:
            n, sh = 4 * (4 * N - i), 1
        phi[p:p + n] = (np.arange(n) + 0.5 * sh) * (2 * np.pi / n)
        p += n
    return phi


class fixed_sim_lib:
    """Simulation library with the duck type filt_simple.library_sepTP needs (get_sim_tmap / get_sim_pmap / hashdict):
    every index returns the maps of one seeded case, scaled by (1 + 0.1 idx)."""

    def __init__(self, case):
        self.c = case

    def hashdict(self):
        return {'fixed_sim_lib': int(self.c['nside']), 'lmax': int(self.c['lmax'])}

    def get_sim_tmap(self, idx):
        return self.c['tmap'] * (1. + 0.1 * max(idx, 0))

    def get_sim_pmap(self, idx):
        f = 1. + 0.1 * max(idx, 0)
        return self.c['qmap'] * f, self.c['umap'] * f


def alm_sample(alm, lmax, lfull=128, stride=41):
    """What the goldens keep of a solution alm: every coefficient with l <= lfull, every stride-th coefficient of the
    whole array, and its auto-spectrum is stored separately (the full arrays are 8 MB each at lmax 1024)."""
    ls = alm_ls(lmax)
    return np.concatenate([alm[ls <= lfull], alm[::stride]])
