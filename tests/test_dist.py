"""m-partitioned transforms: partition arithmetic on the CPU, the full exchange pattern with
simulated ranks on one GPU, and -- when the box has at least two GPUs -- real processes over NCCL + CUDA IPC."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import rand_alm

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("nside,mmax,nranks", [(8, 20, 2), (64, 128, 3), (2048, 2048, 8), (4096, 5000, 8), (4096, 4000, 4)])
def test_partition_is_a_balanced_cover(nside, mmax, nranks):
    from plancklens_b200 import dist_sht
    pair_lo, owner = dist_sht.partition(nside, mmax, nranks)
    assert pair_lo[0] == 0 and pair_lo[-1] == 2 * nside and np.all(np.diff(pair_lo) >= 0)
    assert owner.min() >= 0 and owner.max() < nranks and owner.size == mmax + 1
    # ring pairs are dealt by ring-FFT cost -- the fitted per-ring model of plk_dist_partition: power-of-two rings
    # ~ q log2 q, Bluestein cap rings two M-point FFTs whatever their length -- so every rank gets the same modelled time
    npix_pair = np.array([(8 if ip < 2 * nside - 1 else 4) * (ip + 1 if ip < nside else nside) for ip in range(2 * nside)])
    assert npix_pair.sum() == 12 * nside ** 2
    q = np.array([ip + 1 if ip < nside else nside for ip in range(2 * nside)])
    pow2 = (q & (q - 1)) == 0
    M = 2 ** np.ceil(np.log2(np.maximum(2 * q - 1, 1)))
    lg = lambda v: np.ceil(np.log2(np.maximum(v, 1)))
    c = np.where(q <= 8, 0.02 + 1e-5 * (mmax + 1),
                 np.where(pow2, 0.40 * q * lg(q) / (4096. * 12.), 5.2e-6 * 2. * M * lg(M) + 1.9e-5 * q))
    c[-1] *= 0.5
    cost = np.array([c[pair_lo[r]:pair_lo[r + 1]].sum() for r in range(nranks)])
    if nside >= 2048:
        assert np.max(np.abs(cost / cost.mean() - 1)) < 0.02
        # ranks holding the polar caps get fewer pixels than the ranks of the equatorial belt
        pix = np.array([npix_pair[pair_lo[r]:pair_lo[r + 1]].sum() for r in range(nranks)])
        assert pix[0] < pix[-1]
    # Legendre work per rank ~ sum over owned m of (mmax - m + 1): within 5 % at production sizes
    work = np.array([np.sum(mmax - np.where(owner == q)[0] + 1) for q in range(nranks)], dtype=float)
    if mmax >= 2048:
        assert np.max(np.abs(work / work.mean() - 1)) < 0.05


@pytest.mark.gpu
@pytest.mark.parametrize("nside,lmax,nranks", [(16, 40, 2), (64, 128, 3), (64, 150, 8), (256, 300, 4)])
def test_simulated_ranks_equal_single_gpu(nside, lmax, nranks):
    """Same kernels, same summation order: the m-split result is bit-identical to the single-GPU transform."""
    import torch
    from plancklens_b200 import dist_sht, sht
    rng = np.random.default_rng(nside + nranks)
    plan = sht.get_plan(nside, lmax)
    grp = dist_sht.SimGroup(nside, lmax, nranks)
    # every pixel belongs to exactly one rank
    cover = np.zeros(12 * nside ** 2, dtype=int)
    for r in range(nranks):
        for lo, hi in grp.pixel_ranges(r):
            cover[lo:hi] += 1
    assert np.all(cover == 1)
    a = sht.dev_alm(rand_alm(rng, lmax))
    fl = sht.dev_fl(rng.standard_normal(lmax + 1), lmax)
    assert torch.equal(grp.alm2map(a, fl=fl), plan.alm2map(a, fl=fl))
    for spin in (1, 2, 3):
        g, c = sht.dev_alm(rand_alm(rng, lmax, spin)), sht.dev_alm(rand_alm(rng, lmax, spin))
        got, ref = grp.alm2map_spin(g, c, spin, flg=fl), plan.alm2map_spin(g, c, spin, flg=fl)
        assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1])
    m1 = sht.dev_map(rng.standard_normal(12 * nside ** 2))
    m2 = sht.dev_map(rng.standard_normal(12 * nside ** 2))
    assert torch.equal(grp.map2alm(m1, fl=fl), plan.map2alm(m1, fl=fl))
    for spin in (1, 2):
        got, ref = grp.map2alm_spin(m1, m2, spin, flc=fl), plan.map2alm_spin(m1, m2, spin, flc=fl)
        assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1])


@pytest.mark.gpu
def test_simulated_ranks_qe_p(monkeypatch):
    """The 'p' estimator through qe_device with every transform m-split over 4 simulated ranks: bit-identical to the
    single-GPU estimate evaluated with the same kernels (separate leg-product kernels, PLK_QE_FUSED=0), and equal to
    round-off to the single-GPU default, which evaluates the leg products inside the analysis ring kernel."""
    import torch
    import golden_inputs as gi
    from plancklens_b200 import dist_sht, hp, qest, sht
    q = gi.qe_case()
    cls = q['cls']
    d = sht.dev_alm
    twf = hp.almxfl(q['tlm1'], cls['tt']) + hp.almxfl(q['elm1'], cls['te'])
    ewf = hp.almxfl(q['elm1'], cls['ee']) + hp.almxfl(q['tlm1'], cls['te'])
    bwf = hp.almxfl(q['blm1'], cls['bb'])
    args = [d(x) for x in (q['tlm1'], q['elm1'], q['blm1'], twf, ewf, bwf)]
    fused = qest.qe_device(q['nside'], q['lmax'], q['lmax_qlm']).p(*args)
    monkeypatch.setenv('PLK_QE_FUSED', '0')
    ref = qest.qe_device(q['nside'], q['lmax'], q['lmax_qlm']).p(*args)
    got = qest.qe_device(q['nside'], q['lmax'], q['lmax_qlm'],
                         plan_ivf=dist_sht.SimGroup(q['nside'], q['lmax'], 4),
                         plan_qlm=dist_sht.SimGroup(q['nside'], q['lmax_qlm'], 4)).p(*args)
    assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1])
    for a, b in zip(fused, ref):
        assert float(torch.linalg.norm(a - b) / torch.linalg.norm(b)) < 1e-13


@pytest.mark.gpu
def test_real_processes_two_gpus():
    """torchrun x 2 over NCCL + CUDA IPC peer stores; skipped on single-GPU boxes."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(n),
           '--master-addr', '127.0.0.1', '--master-port', '29631', os.path.join(HERE, 'dist_worker.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert 'DIST OK' in r.stdout


@pytest.mark.gpu
def test_m_distributed_cg_two_gpus():
    """qcinv/dist_cg.py: masked-sky T and P filters (default chains, nside 512 / lmax 1024) with the forward operator
    split by m over real processes -- same iteration count and eps trace as the single-GPU solve, solution to 1e-8;
    skipped on single-GPU boxes."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(n),
           '--master-addr', '127.0.0.1', '--master-port', '29633', os.path.join(HERE, 'dist_cg_worker.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert 'DIST CG OK' in r.stdout
