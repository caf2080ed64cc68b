"""Generates tests/golden/reference_golden_wrappers.npz with the UNMODIFIED reference (healpy shim on the path): the
small host-side wrappers around the hot path -- filt_util.library_fml / _alm_copy, filt_simple.library_fullsky_alms_sepTP,
sims.utils.sim_lib_add_sim / add_dat, sims.maps.cmb_maps_harmonicspace, utils.stats, utils.alm2rlm.
Run from the repo root:  python tests/golden/make_golden_wrappers.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'healpy_shim'))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import healpy as hp  # the shim  # noqa: E402
from plancklens import utils  # noqa: E402  (reference)
from plancklens.filt import filt_simple, filt_util  # noqa: E402
from plancklens.sims import maps, utils as sims_utils  # noqa: E402

import golden_inputs as gi  # noqa: E402

q = gi.qe_case()
w = gi.wrapper_case(q)
out = {}
iv = gi.idx_ivfs(q, hp)
iv.get_ftl = lambda: w['ftl']
iv.get_fel = lambda: w['fel']
iv.get_fbl = lambda: w['fbl']
fml = filt_util.library_fml(iv, w['lmax_cut'], w['fm_t'], w['fm_e'], w['fm_b'])
for name in ('get_ftl', 'get_fel', 'get_fbl'):
    out['fml_' + name] = getattr(fml, name)()
for name in ('tlm', 'elm', 'blm', 'tmliklm', 'emliklm', 'bmliklm'):
    out['fml_' + name] = getattr(fml, 'get_sim_' + name)(2)
out['alm_copy_up'] = filt_util._alm_copy(q['tlm1'], None, q['lmax'] + 5, q['lmax'] + 2)
out['alm_copy_dn'] = filt_util._alm_copy(q['tlm1'], -1, q['lmax'] - 7, 9)

sims = gi.alm_sims(q)
with tempfile.TemporaryDirectory() as tmp:
    lib = filt_simple.library_fullsky_alms_sepTP(os.path.join(tmp, 'f'), sims, {'t': w['transf'], 'e': w['transf'], 'b': w['transf'] ** 2},
                                                 q['cls'], w['ftl'], w['fel'], w['fbl'], cache=False)
    for name in ('tlm', 'elm', 'blm', 'tmliklm', 'emliklm'):
        out['alms_' + name] = getattr(lib, 'get_sim_' + name)(1)
    out['alms_tal_b'] = lib.get_tal('b')

add_sim = sims_utils.sim_lib_add_sim([gi.map_sims(q, 1.0), gi.map_sims(q, -0.3)], weights=[0.7, 2.0])
add_dat = sims_utils.sim_lib_add_dat([gi.map_sims(q, 1.0), gi.map_sims(q, -0.3)])
for tag, lib in (('add_sim', add_sim), ('add_dat', add_dat)):
    for idx in (-1, 2):
        out['%s_t_%d' % (tag, idx)] = lib.get_sim_tmap(idx)
        out['%s_q_%d' % (tag, idx)], out['%s_u_%d' % (tag, idx)] = lib.get_sim_pmap(idx)

hs = maps.cmb_maps_harmonicspace(sims, {'t': w['transf'], 'e': w['transf'], 'b': w['transf']},
                                 {'t': w['nl_t'], 'e': w['nl_p'], 'b': w['nl_p']}, gi.fixed_phas(q))
out['hs_tlm'] = hs.get_sim_tmap(1)
out['hs_elm'], out['hs_blm'] = hs.get_sim_pmap(1)
hsm = maps.cmb_maps_harmonicspace(sims, {'t': w['transf'], 'e': w['transf'], 'b': w['transf']},
                                  {'t': w['nl_t'], 'e': w['nl_p'], 'b': w['nl_p']}, gi.fixed_phas(q), nside=8)
out['hs_tmap'] = hsm.get_sim_tmap(1)
out['hs_qmap'], out['hs_umap'] = hsm.get_sim_pmap(1)

rng = np.random.default_rng(5)
st = utils.stats(6)
rows = rng.standard_normal((20, 6)) * np.arange(1, 7)
for r in rows:
    st.add(r)
out['stats_rows'] = rows
out['stats_mean'], out['stats_cov'], out['stats_sig'] = st.mean(), st.cov(), st.sigmas()
out['stats_som'], out['stats_corr'], out['stats_inv'] = st.sigmas_on_mean(), st.corrcoeffs(), st.inverse()
out['stats_chisq'] = np.array([st.get_chisq(rows[3] * 0.5), st.get_chisq_pte(rows[3] * 0.5)])
rb = st.rebin_that_nooverlap(np.arange(6.), np.array([0, 2, 4]), np.array([1, 3, 5]), weights=np.arange(1., 7.))
out['stats_rb_mean'], out['stats_rb_cov'] = rb.mean(), rb.cov()
out['rlm'] = utils.alm2rlm(q['tlm1'])
fn = os.path.join(ROOT, 'tests', 'golden', 'reference_golden_wrappers.npz')
np.savez_compressed(fn, **out)
print('wrote', fn, len(out), 'arrays')
