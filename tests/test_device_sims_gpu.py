"""GPU: counter-based random numbers (Philox4x32-10 kernels) against the numpy oracle -- bit exact for the generator
words, 1e-13 for the Gaussian mapping (libm vs CUDA log / sincospi) -- and the device-resident simulation -> filter -> QE
path against the host-fed path of the same libraries."""
import os
import tempfile

import numpy as np
import pytest

from helpers import rel_l2

pytestmark = pytest.mark.gpu


def test_philox_words_bit_exact():
    from oracle import ref_rng
    from plancklens_b200 import sht
    for seed, stream, n in ((0, 0, 7), (20000, (5 << 8) | 2, 100003), ((0xdeadbeef << 32) | 0x1234, (1 << 40) + 17, 4099)):
        got = sht.philox_words(seed, stream, n).cpu().numpy().view(np.uint32)
        assert np.array_equal(got, ref_rng.words(seed, stream, n))


@pytest.mark.parametrize("n", [1, 2, 5, 100001])
def test_randn_matches_oracle(n):
    import torch
    from oracle import ref_rng
    from plancklens_b200 import sht
    z = sht.randn(20000, 77, n).cpu().numpy()
    ref = ref_rng.randn(20000, 77, n)
    assert np.max(np.abs(z - ref)) < 1e-13 * max(1.0, np.max(np.abs(ref)))
    base = torch.arange(n, dtype=torch.float64, device='cuda')
    z2 = sht.randn(20000, 77, n, scale=0.25, add=base).cpu().numpy()
    assert np.max(np.abs(z2 - (np.arange(n) + 0.25 * ref))) < 1e-12 * max(1.0, n)


def test_randn_alm_matches_oracle_and_recipe():
    from oracle import ref_rng
    from plancklens_b200 import sht
    from plancklens_b200.sims import phas
    lmax = 200
    a = sht.randn_alm(10000, 3, lmax).cpu().numpy()
    assert np.max(np.abs(a - ref_rng.randn_alm(10000, 3, lmax))) < 1e-13
    assert np.all(a[:lmax + 1].imag == 0)
    lib = phas.lib_phas(None, 3, lmax, device=True)
    assert lib.hashdict()['rng'] == 'philox4x32-10'
    assert np.array_equal(lib.get_sim(4, idf=1), lib.get_sim_dev(4, 1).cpu().numpy())
    assert np.array_equal(lib.get_sim(4, idf=1), ref_rng.randn_alm(10000, phas._stream_id(4, 1), lmax)) or \
        np.max(np.abs(lib.get_sim(4, idf=1) - ref_rng.randn_alm(10000, phas._stream_id(4, 1), lmax))) < 1e-13
    pix = phas.pix_lib_phas(None, 3, (12 * 16 ** 2,), device=True)
    assert pix.get_sim(-1, idf=0).shape == (12 * 16 ** 2,)
    assert not np.array_equal(pix.get_sim(-1, idf=0), pix.get_sim(0, idf=0))
    big = sht.randn(1, 2, 4000000).cpu().numpy()
    assert abs(big.mean()) < 3e-3 and abs(big.std() - 1) < 3e-3 and abs(np.mean(big ** 4) - 3) < 0.03


def test_device_resident_pipeline_equals_host_fed_pipeline(oracle_sht):
    """params/idealized_example.py with PLK_DEVICE_SIMS=1: simulated maps, filtered alms and the 'p' estimate stay on the
    GPU (`get_sim_*map_dev` -> `get_sim_teblm_dev` -> `get_sim_qlm_dev`).  Same numbers as feeding the libraries' host
    accessors through the oracle, and as the cached host API of the same library."""
    from oracle import ref_qe
    from oracle.healpy_shim.healpy import almxfl
    from test_pipeline_gpu import _load_params
    with tempfile.TemporaryDirectory() as tmp:
        par = _load_params('idealized_example', {'PLENS': tmp, 'PLK_NSIDE': '64', 'PLK_LMAX_IVF': '96', 'PLK_LMAX_QLM': '128',
                                                 'PLK_NSIMS': '4', 'PLK_DEVICE_SIMS': '1'})
        assert par.qlms_dd._dev_ok()
        Gd, Cd = par.qlms_dd.get_sim_qlm_dev('p', 0)
        G = par.qlms_dd.get_sim_qlm('p', 0)            # host API of the same library (cached on disk)
        assert rel_l2(Gd.cpu().numpy(), G) < 1e-12
        assert os.path.exists(os.path.join(par.qlms_dd.lib_dir, 'sim_p_0000.fits'))
        # host accessors of the simulation library return the very numbers the device path used
        tmap = par.sims.get_sim_tmap(0)
        q, u = par.sims.get_sim_pmap(0)
        fac = np.where(par.transf > 0, 1 / par.transf, 0)
        t = almxfl(oracle_sht.map2alm(tmap, lmax=par.lmax_ivf), par.ftl * fac)
        e, b = oracle_sht.map2alm_spin([q, u], 2, lmax=par.lmax_ivf)
        e, b = almxfl(e, par.fel * fac), almxfl(b, par.fbl * fac)
        td, ed, bd = [x.cpu().numpy() for x in par.ivfs.get_sim_teblm_dev(0)]
        assert rel_l2(td, t) < 1e-10 and rel_l2(ed, e) < 1e-10 and rel_l2(bd, b) < 1e-10
        cls = {k: par.cl_len[k] for k in ['tt', 'ee', 'bb', 'te']}
        Gr, Cr = ref_qe.qe('p', t, e, b, cls, par.nside, par.lmax_qlm)
        assert rel_l2(Gd.cpu().numpy(), Gr) < 1e-10 and rel_l2(Cd.cpu().numpy(), Cr) < 1e-10
        # sim x data library (different legs -> symmetrised), device vs host API
        Gs, _ = par.qlms_ds.get_sim_qlm_dev('p_p', 1)
        assert rel_l2(Gs.cpu().numpy(), par.qlms_ds.get_sim_qlm('p_p', 1)) < 1e-12
        # noise level of the simulated maps: white noise of nlev_t uK-arcmin on top of the CMB
        n1 = par.sims.get_sim_tmap(1) - par.sims.get_sim_tmap(0)
        assert np.std(n1) > par.nlev_t / (np.sqrt(4 * np.pi / tmap.size) * 180 * 60 / np.pi)


@pytest.mark.parametrize("nside_in,nside_out", [(64, 64), (64, 16), (256, 32), (2048, 128)])
def test_device_ud_grade_matches_oracle(nside_in, nside_out):
    """plk_udgrade_sum_dev (hp.ud_grade(power=-2) of the coarse multigrid levels, opfilt_tt.py:172-181) against the
    oracle's NEST-based sum of children -- bit for bit (see hpx_children_sum in plk_blas.cuh for why that matters)"""
    from oracle import ref_geom as rg
    from plancklens_b200 import sht
    rng = np.random.default_rng(nside_in)
    m = rng.standard_normal(12 * nside_in ** 2)
    got = sht.ud_grade_sum(sht.dev_map(m), nside_out).cpu().numpy()
    ref = rg.ud_grade_sum(m, nside_out)
    assert np.array_equal(got, ref)       # children added in numpy's order: bit-identical to healpy-style ud_grade
