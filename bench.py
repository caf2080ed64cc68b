#!/usr/bin/env python
"""Benchmark of the plancklens hot path on B200.

Contract (driver): `python bench.py --gpus N --steps K --warmup W` (under torchrun for N > 1) prints ONE JSON line.

Headline workload (BASELINE.json configs[1]): idealized full-sky MV 'p' lensing QE, nside 2048, lmax_ivf = lmax_qlm = 2048,
synthetic Gaussian CMB + white-noise skies filtered isotropically (params/idealized_example.py of the reference with the
FFP10 fiducial spectra, 5' beam, 35 / 55 uK-arcmin).  One step = `qest.library.get_sim_qlm('p', idx)` for one simulation
starting from its cached inverse-variance filtered alms (what `run_qlms.py -k p -dd` does after `-ivt -ivp`): 1 spin-0 +
4 spin-s syntheses, per-pixel products, spin-1 analysis -> (glm, clm).
  value    : steps/s with the filtered alms resident in HBM (CUDA events, max over ranks)
  e2e      : the same through `qest.library.eval_qlms` with numpy (pinned host) inputs and numpy outputs, every H2D + D2H
             inside the timed region
  roofline : the dominant kernel, `legendre_synth_kernel<spin>`; bound = FP64 FMA pipe.  `achieved` / `frac` count the DFMAs
             the kernel EXECUTES (24 flop x the (l, m, ring pair) volume it walks) against the nominal FP64 FMA peak at the
             sampled SM clock; the full-volume figure of SURVEY.md section 8d is given beside it.  `kernels` holds the same for
             the analysis kernels and the HBM roofline of the ring-FFT stage against MEASURED_PEAKS.json
  extra.target : the north_star target -- masked-sky CG-filtered 'p' QE (cinv_t + cinv_p with the reference's default
             multigrid chains, eps 1e-5, synthetic Galactic mask + point-source holes + anisotropic noise) at nside 2048,
             lmax 2048 and lmax 3000, simulations drawn, filtered and estimated on the GPU, sharded idx % N over ranks:
             simulations/s, CG iterations and iterations/s, per-stage ms, algorithmic TFLOP/s and roofline fraction
  extra.dist   : (N > 1) BASELINE.json configs[4]: ONE 'p' estimate at nside 4096 / lmax_ivf 4000 / lmax_qlm 5000 with every
             transform m-partitioned over the N GPUs, checked bit-for-bit against the single-GPU plan on rank 0;
             extra.target.lmax2048.dist_cg: the masked T and P filters of one simulation with the forward operator
             m-partitioned (dots as scalar all-reduces), against the single-GPU solve of the same maps
  cpu_baseline : the CPU oracle port of the headline step on a bounded sample, host cores stated
`--impl reference` runs the CPU port alone: the FULL step (every m), for as many of the requested steps as fit a time
budget, and prints the ms_per_step it measured (the reference itself needs healpy, which is absent here).
Multi-GPU: simulations are sharded over ranks (idx % N == rank, reference: examples/run_qlms.py:72), weak scaling, one NCCL
reduce of the accumulated qlm (the mean-field sum of qest.py:239-243) inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NSIDE, LMAX_IVF, LMAX_QLM = 2048, 2048, 2048
NLEV_T, NLEV_P, BEAM_AMIN, LMIN_IVF = 35., 55., 5., 100
NPOOL = 3
TARGET_LMAX = (2048, 3000)
DIST_CFG = (4096, 4000, 5000)      # nside, lmax_ivf, lmax_qlm of BASELINE.json configs[4]


def alm_size(lmax):
    return (lmax + 1) * (lmax + 2) // 2


def n_lm(lmax):
    return sum(lmax - m + 1 for m in range(lmax + 1))


def F0(lmax, nside):
    """algorithmic flop of one spin-0 Legendre stage (8 per (l, m, ring pair))"""
    return 8.0 * n_lm(lmax) * 2 * nside


def Fs(lmax, nside):
    """same for spin s > 0: 24 per (l, m, ring pair)"""
    return 24.0 * n_lm(lmax) * 2 * nside


def fiducial(lmax, beam_amin=None, nlev_t=None, nlev_p=None):
    from plancklens_b200 import hp, utils
    beam_amin = BEAM_AMIN if beam_amin is None else beam_amin
    nlev_t = NLEV_T if nlev_t is None else nlev_t
    nlev_p = NLEV_P if nlev_p is None else nlev_p
    cls = utils.camb_clfile(os.path.join(ROOT, 'plancklens_b200', 'data', 'cls', 'FFP10_wdipole_lensedCls.dat'), lmax=lmax)
    transf = hp.gauss_beam(beam_amin / 60. / 180. * np.pi, lmax=lmax)
    ftl = utils.cli(cls['tt'] + (nlev_t / 60. / 180. * np.pi / transf) ** 2)
    fel = utils.cli(cls['ee'] + (nlev_p / 60. / 180. * np.pi / transf) ** 2)
    fbl = utils.cli(cls['bb'] + (nlev_p / 60. / 180. * np.pi / transf) ** 2)
    for f in (ftl, fel, fbl):
        f[:LMIN_IVF] = 0.
    return cls, transf, ftl, fel, fbl


def filtered_sim(idx, lmax, cls, transf, fls, nlev_t=None, nlev_p=None):
    """Inverse-variance filtered alms of one synthetic sky: f_l (a_lm + n_lm / b_l), correlated T/E draw
    (reference recipe: sims/phas.py:162-168, sims/cmbs.py:35-69, white noise of sims/maps.py:136-173 in harmonic space)."""
    from plancklens_b200 import hp
    nlev_t = NLEV_T if nlev_t is None else nlev_t
    nlev_p = NLEV_P if nlev_p is None else nlev_p
    rng = np.random.default_rng(10000 + idx)
    n = alm_size(lmax)

    def phase():
        a = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) / np.sqrt(2.)
        a[:lmax + 1] = np.sqrt(2.) * a[:lmax + 1].real
        return a
    p1, p2, p3 = phase(), phase(), phase()
    tt, ee, bb, te = cls['tt'], cls['ee'], cls['bb'], cls['te']
    st = np.sqrt(np.maximum(tt, 0))
    r = np.where(st > 0, te / np.where(st > 0, st, 1), 0.)
    tlm = hp.almxfl(p1, st)
    elm = hp.almxfl(p1, r) + hp.almxfl(p2, np.sqrt(np.maximum(ee - r ** 2, 0)))
    blm = hp.almxfl(p3, np.sqrt(np.maximum(bb, 0)))
    bi = np.where(transf > 0, 1. / transf, 0.)
    out = []
    for alm, nlev, fl in ((tlm, nlev_t, fls[0]), (elm, nlev_p, fls[1]), (blm, nlev_p, fls[2])):
        noise = hp.almxfl(phase(), (nlev / 60. / 180. * np.pi) * bi)
        out.append(hp.almxfl(alm + noise, fl))
    return out


class mem_ivfs:
    """In-memory filtering library (pinned host arrays) with the duck type `qest.library` needs."""

    def __init__(self, sims, cl, nside):
        self.sims, self.cl, self.nside, self.lib_dir = sims, cl, nside, None

    def hashdict(self):
        return {'bench': len(self.sims)}

    def get_fmask(self):
        return np.ones(1)

    def get_sim_tlm(self, idx): return self.sims[idx % len(self.sims)][0]
    def get_sim_elm(self, idx): return self.sims[idx % len(self.sims)][1]
    def get_sim_blm(self, idx): return self.sims[idx % len(self.sims)][2]

    def get_sim_tmliklm(self, idx):
        from plancklens_b200 import hp
        return hp.almxfl(self.get_sim_tlm(idx), self.cl['tt'])

    def get_sim_emliklm(self, idx):
        from plancklens_b200 import hp
        return hp.almxfl(self.get_sim_elm(idx), self.cl['ee'])

    def get_sim_bmliklm(self, idx):
        from plancklens_b200 import hp
        return hp.almxfl(self.get_sim_blm(idx), self.cl['bb'])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace('.', '').isdigit()]
        reasons = []
        for name, col in (('hw_slowdown', 4), ('hw_thermal_slowdown', 5), ('sw_thermal_slowdown', 6), ('sw_power_cap', 7)):
            if any(len(r) >= 8 and r[col].lower().startswith('active') for r in self.rows):
                reasons.append(name)
        pw = [float(r[3]) for r in self.rows if len(r) >= 8 and r[3].replace('.', '').isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return json.load(f)
    except Exception:
        return {}


def ncu_traffic(kernel_substr):
    """dram bytes read + written per launch of a kernel, from the ncu summary kept under profiles/ (None if absent)"""
    for name in ('r02_ncu_metrics.json', 'r01_ncu_metrics.json'):
        fn = os.path.join(ROOT, 'profiles', name)
        if os.path.exists(fn):
            try:
                with open(fn) as f:
                    d = json.load(f)
                for k, v in d.get('kernels', {}).items():
                    if kernel_substr in k and v.get('dram_read_bytes') is not None:
                        return float(v['dram_read_bytes']) + float(v['dram_write_bytes']), 'profiles/' + name + ': ' + k
            except Exception:
                pass
    return None, None


# ------------------------------------------------------------------------------------------------ CPU port
def cpu_port_step(sims, cls, mstep, workers=-1):
    """One 'p' estimate with the CPU oracle (oracle/csht.c Legendre stage with OpenMP, scipy ring FFTs on `workers`
    threads, numpy pixel products).  mstep = 1 runs the complete step; mstep > 1 runs the Legendre stages on every
    mstep-th m and scales that part by the sampled share of the (l, m) work (bounded sample for `cpu_baseline`).
    Returns (seconds for the full step -- measured if mstep == 1, estimated otherwise --, description)."""
    from oracle import ref_sht
    from oracle.healpy_shim.healpy import almxfl
    tbar, ebar, bbar = sims
    lmax = LMAX_IVF
    l = np.arange(lmax + 1, dtype=float)
    work_all = sum(lmax - m + 1 for m in range(lmax + 1))
    work_s = sum(lmax - m + 1 for m in range(0, lmax + 1, mstep))
    scale = work_all / work_s
    twf = almxfl(tbar, cls['tt']) + almxfl(ebar, cls['te'])
    ewf = almxfl(ebar, cls['ee']) + almxfl(tbar, cls['te'])
    bwf = almxfl(bbar, cls['bb'])
    f3 = np.sqrt(np.maximum((l - 2) * (l + 3), 0)); f3[:3] = 0
    f1 = np.sqrt(np.maximum((l + 2) * (l - 1), 0)); f1[:1] = 0
    t_leg = t_fft = t_pix = 0.0

    def synth(spin, a, b=None):
        nonlocal t_leg, t_fft
        t0 = time.time()
        X1, X2 = ref_sht.legendre_synth(NSIDE, spin, lmax, lmax, a, b, mstep=mstep)
        t_leg += time.time() - t0
        t0 = time.time()
        out = [ref_sht.phase2map(NSIDE, X, workers=workers) for X in ((X1, X2) if spin else (X1,))]
        t_fft += time.time() - t0
        return out
    (t,) = synth(0, tbar)
    g1t, c1t = synth(1, almxfl(twf, -np.sqrt(l * (l + 1))), np.zeros_like(twf))
    q, u = synth(2, 0.5 * ebar, 0.5 * bbar)
    g3, c3 = synth(3, almxfl(ewf, f3), almxfl(bwf, f3))
    g1, c1 = synth(1, almxfl(ewf, f1), almxfl(bwf, f1))
    t0 = time.time()
    re_t, im_t = g1t * t, c1t * t                                    # qest.py:256-257
    re_p = (q * g3 + u * c3) - (q * g1 + u * c1)                     # qest.py:276-278
    im_p = (q * c3 - u * g3) - (u * g1 - q * c1)
    t_pix += time.time() - t0
    for re, im in ((re_t, im_t), (re_p, im_p)):                      # the reference runs two analyses and sums the qlm
        t0 = time.time()
        X1, X2 = ref_sht.map2phase(NSIDE, re, LMAX_QLM, workers=workers), ref_sht.map2phase(NSIDE, im, LMAX_QLM, workers=workers)
        t_fft += time.time() - t0
        t0 = time.time()
        ref_sht.legendre_anal(NSIDE, 1, LMAX_QLM, LMAX_QLM, X1, X2, mstep=mstep)
        t_leg += time.time() - t0
    total = t_leg * scale + t_fft + t_pix
    if mstep == 1:
        desc = "complete step: Legendre %.1f s, ring FFTs %.1f s, pixel products %.1f s" % (t_leg, t_fft, t_pix)
    else:
        desc = "Legendre stages on every %d-th m (%.1f s x %.1f), ring FFTs of all 9 + 4 components (%.1f s) and pixel " \
               "products (%.1f s) in full" % (mstep, t_leg, scale, t_fft, t_pix)
    return total, desc


def cpu_cg_iteration(cls, transf, nside=2048, lmax=2048):
    """One top-level forward operator of the masked temperature filter with the CPU oracle (opfilt_tt.fwd_op: two spin-0
    transforms + N^-1 with monopole/dipole projection), seconds -- the CPU leg of the CG half of the metric."""
    from oracle import ref_cg
    rng = np.random.default_rng(3)
    npix = 12 * nside ** 2
    ninv = np.ones(npix)
    ninv[npix // 3:2 * npix // 3] = 0.
    nf = ref_cg.ninv_tt(ninv, transf, marge_monopole=True, marge_dipole=True)
    x = (rng.standard_normal(alm_size(lmax)) + 1j * rng.standard_normal(alm_size(lmax)))
    x[:lmax + 1] = x[:lmax + 1].real
    t0 = time.time()
    ref_cg.fwd_tt(x, cls['tt'], nf)
    return time.time() - t0


def run_reference(args):
    """--impl reference: the CPU port of the step (healpy is not installable here), rank 0 only.  Every step is the
    COMPLETE 'p' estimate (all m); steps are capped by a wall-clock budget and the line reports the steps actually run."""
    if int(os.environ.get('RANK', 0)) != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm runs on rank 0 alone and gets all host cores
    os.environ['OMP_NUM_THREADS'] = str(os.cpu_count())
    from oracle import ref_sht
    ref_sht.build()
    cls, transf, ftl, fel, fbl = fiducial(LMAX_IVF)
    sims = filtered_sim(0, LMAX_IVF, cls, transf, (ftl, fel, fbl))
    cores = ref_sht.max_threads()
    budget = float(args.ref_budget_s)
    t_start = time.time()
    ts, desc = [], ''
    nwarm = 0
    # one warm-up step when the budget allows (first-touch of the FFT plans and page faults of the 400 MB maps)
    t, desc = cpu_port_step(sims, cls, 1)
    if 2.5 * t < budget and args.warmup > 0:
        nwarm = 1
    else:
        ts.append(t)
    while len(ts) < args.steps and (time.time() - t_start) + (ts[-1] if ts else t) < budget:
        t, desc = cpu_port_step(sims, cls, 1)
        ts.append(t)
    t = float(np.mean(ts))
    t_cg = cpu_cg_iteration(cls, transf)
    line = {"impl": "reference", "metric": "QE qlms/sec ('p', nside 2048, lmax 2048)", "value": 1.0 / t, "unit": "qlm/s",
            "n_gpus": args.gpus, "steps": len(ts), "warmup": nwarm, "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": 1e3 * t, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(),
            "cpu_baseline": {"value": 1.0 / t, "unit": "qlm/s", "cores": cores, "kind": "port",
                             "sample": "%d complete steps, every m (%s); wall-clock budget %.0f s" % (len(ts), desc, budget)},
            "e2e": {"value": 1.0 / t, "unit": "qlm/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "extra": {"masked_cg_T_fwd_op_s": t_cg, "masked_cg_T_note": "one top-level opfilt_tt forward operator (2 spin-0 "
                      "transforms + N^-1 + monopole/dipole projection) at nside 2048 / lmax 2048 on the CPU port; a CG-T "
                      "iteration of the default chain costs this plus the multigrid preconditioner"},
            "note": "CPU oracle port (oracle/csht.c with OpenMP + threaded scipy ring FFTs): the reference's own path needs "
                    "healpy, which is not installed and not installable offline.  Steps are complete (not sampled); the run "
                    "stops at the wall-clock budget, so `steps` may be fewer than requested"}
    print(json.dumps(line))


def workload_config():
    return {"workload": "idealized full-sky MV 'p' lensing QE from cached inverse-variance filtered alms "
                        "(BASELINE.json configs[1]): nside 2048, lmax_ivf 2048, lmax_qlm 2048, 1 spin-0 + 4 spin-s "
                        "syntheses + spin-1 analysis per estimate",
            "nside": NSIDE, "lmax_ivf": LMAX_IVF, "lmax_qlm": LMAX_QLM, "key": "p",
            "l2_policy": "every step streams ~6 GB of maps and phase arrays (>> 126 MB L2); inputs rotate over %d sims" % NPOOL,
            "parallelism": "simulations sharded over ranks (idx % N == rank)"}


# ------------------------------------------------------------------------------------------------ north_star target
def synthetic_sky_model(nside):
    """Mask and anisotropic noise shape of SURVEY.md section 8d (|b| < 20 deg + 2000 discs of 10'; 1 + 0.5 z^2)."""
    sys.path.insert(0, os.path.join(ROOT, 'scripts'))
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import bench_cg
    return bench_cg.synthetic_mask(nside, np.random.default_rng(7))


def build_target(lmax, tmp, mask, z):
    """The libraries of params/anisofilt_example.py (reference params/anisofilt_example.py:62-104) for the synthetic
    masked sky: device-drawn sims -> cinv_t / cinv_p (default chains) -> library_cinv_sepTP -> library_ftl -> qlms_dd."""
    from plancklens_b200 import hp, qest
    from plancklens_b200.filt import filt_cinv, filt_util
    from plancklens_b200.sims import cmbs, maps, phas
    cls, transf, _, _, _ = fiducial(lmax)
    cl_ivf = {k: cls[k][:lmax + 1] for k in ('tt', 'ee', 'bb', 'te')}
    vamin = np.sqrt(hp.nside2pixarea(NSIDE, degrees=True)) * 60
    ninv_t = mask * (vamin / NLEV_T) ** 2 * (1 + 0.5 * z ** 2)
    ninv_p = mask * (vamin / NLEV_P) ** 2 * (1 + 0.5 * z ** 2)
    pix_phas = phas.pix_lib_phas(None, 3, (hp.nside2npix(NSIDE),), device=True)
    cmb_sims = cmbs.sims_cmb_unl(cl_ivf, phas.lib_phas(None, 3, lmax, device=True))
    sims = maps.cmb_maps_nlev(cmb_sims, transf, NLEV_T, NLEV_P, NSIDE, pix_lib_phas=pix_phas)
    d = os.path.join(tmp, 'lmax%d' % lmax)
    cinv_t = filt_cinv.cinv_t(os.path.join(d, 'cinv_t'), lmax, NSIDE, cl_ivf, transf, [ninv_t], marge_monopole=True,
                              marge_dipole=True, marge_maps=[])
    cinv_p = filt_cinv.cinv_p(os.path.join(d, 'cinv_p'), lmax, NSIDE, cl_ivf, transf, [[ninv_p]])
    ivfs_raw = filt_cinv.library_cinv_sepTP(os.path.join(d, 'ivfs'), sims, cinv_t, cinv_p, cls)
    cut = np.ones(lmax + 1) * (np.arange(lmax + 1) >= LMIN_IVF)
    ivfs = filt_util.library_ftl(ivfs_raw, lmax, cut, cut, cut)
    qlms_dd = qest.library_sepTP(os.path.join(d, 'qlms_dd'), ivfs, ivfs, cls['te'], NSIDE, lmax_qlm=lmax)
    return {'sims': sims, 'cinv_t': cinv_t, 'cinv_p': cinv_p, 'ivfs_raw': ivfs_raw, 'ivfs': ivfs, 'qlms_dd': qlms_dd}


def cg_flops(lmax, it_t, it_p, algorithmic=True):
    """Legendre flop of one masked T + P filtering with the default chains: per top-level iteration one forward operator
    at full resolution plus one multigrid preconditioner, plus calc_prep (one analysis) per solve.
    algorithmic=True : the count of SURVEY.md section 8d, i.e. the work of the REFERENCE's loop (cd_solve.py:61-102 applies
                       the preconditioner once more per stage than its result needs: 4 applications for 3 iterations):
                       CG-T 2 F0 + 6 F0(1024,512) + 24 F0(512,256) + 96 F0(256,128), CG-P 2 Fs + 8 Fs(1024,512) + 32 Fs(512,256)
    algorithmic=False: what this implementation executes (3 applications per stage: 6 / 18 / 54 and 6 / 18)."""
    if algorithmic:
        pre_t = 6 * F0(1024, 512) + 24 * F0(512, 256) + 96 * F0(256, 128)
        pre_p = 8 * Fs(1024, 512) + 32 * Fs(512, 256)
    else:
        pre_t = 6 * F0(1024, 512) + 18 * F0(512, 256) + 54 * F0(256, 128)
        pre_p = 6 * Fs(1024, 512) + 18 * Fs(512, 256)
    ft = F0(lmax, NSIDE) + it_t * (2 * F0(lmax, NSIDE) + pre_t)
    fp = Fs(lmax, NSIDE) + it_p * (2 * Fs(lmax, NSIDE) + pre_p)
    return ft, fp


def run_dist_cg(lib, lmax, rank, world, dist):
    """The masked-sky T and P filters of one simulation with the forward operator m-partitioned over the ranks
    (qcinv/dist_cg.py), next to the single-GPU solve of the same maps: iterations, ms, speed-up."""
    import torch
    from plancklens_b200.qcinv import dist_cg, util_alm
    idx = 7777                                   # the same simulation on every rank (counter-based draws)
    tmap = lib['sims'].get_sim_tmap_dev(idx)
    qmap, umap = lib['sims'].get_sim_pmap_dev(idx)
    out = {}
    for name, cinv, maps, zero in (
            ('T', lib['cinv_t'], tmap, lambda: util_alm.dalm.zeros(lmax)),
            ('P', lib['cinv_p'], [qmap, umap], lambda: util_alm.eblm([util_alm.dalm.zeros(lmax), util_alm.dalm.zeros(lmax)]))):
        chain = cinv.chain
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ref = zero()
        torch.cuda.synchronize()
        e[0].record()
        chain.solve(ref, maps)
        e[1].record()
        n1 = chain.niter
        dc = dist_cg.dist_chain(chain)
        dc.solve(zero(), maps)                   # warm-up of the m-partitioned plan
        dist.barrier()
        got = zero()
        e[2].record()
        n = dc.solve(got, maps)
        e[3].record()
        torch.cuda.synchronize()
        t = torch.tensor([e[0].elapsed_time(e[1]), e[2].elapsed_time(e[3])], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        err = max(float(torch.linalg.norm(a.t - b.t) / torch.linalg.norm(b.t)) for a, b in zip(dist_cg._comps(got), dist_cg._comps(ref)))
        out[name] = {"iterations_single_gpu": int(n1), "iterations_m_partitioned": int(n), "ms_single_gpu": float(t[0].item()),
                     "ms_m_partitioned": float(t[1].item()), "speedup": float(t[0].item() / t[1].item()),
                     "iter_per_s_m_partitioned": n / (float(t[1].item()) * 1e-3), "rel_l2_vs_single_gpu": err}
        del dc
    out["note"] = "forward operator split by m over the ranks, dots as scalar all-reduces, multigrid preconditioner replicated " \
                  "on every rank (Amdahl: only the two full-resolution transforms of an iteration are split)"
    return out


def run_target(lmax, nsims, tmp, mask, z, rank, world, dist, peak_nominal, with_dist_cg=False):
    import torch
    from plancklens_b200 import sht
    t0 = time.perf_counter()
    lib = build_target(lmax, tmp, mask, z)
    q = lib['qlms_dd']
    _ = lib['cinv_t'].chain.bstage, lib['cinv_p'].chain.bstage       # degraded filters, dense preconditioners, stages
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0
    nalm = alm_size(lmax)
    mf = [torch.zeros(nalm, dtype=torch.complex128, device='cuda') for _ in range(2)]

    def one(idx, prefetch=()):
        G, C = q.get_sim_qlm_dev('p', idx, prefetch=prefetch)
        sht.alm_axpy(mf[0], G, 1.0)
        sht.alm_axpy(mf[1], C, 1.0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    # two warm-up simulations per rank: plans and tables, then the CUDA-graph capture of the preconditioners
    one(1000 * rank + 0)
    one(1000 * rank + 1)
    barrier()
    its_t, its_p = [], []
    n0 = sht._lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw = time.perf_counter()
    e0.record()
    idxs = [1000 * rank + 2 + i for i in range(nsims)]     # distinct simulations on every rank: idx % N sharding of a batch
    depth = int(os.environ.get('PLK_TP_PREFETCH', '2'))
    for i, idx in enumerate(idxs):
        # the loop knows which simulations come next: their filters run while this estimate is evaluated (all of it
        # inside the timed region: nothing of a timed simulation is started before e0)
        one(idx, prefetch=idxs[i + 1:i + 1 + depth])
        it = lib['ivfs_raw'].cg_iterations[idx]
        its_t.append(int(it['T']))
        its_p.append(int(it['P']))
    if world > 1:
        for v in mf:
            dist.reduce(torch.view_as_real(v), dst=0)
    e1.record()
    barrier()
    wall = time.perf_counter() - tw
    lib['ivfs'].flush()
    launches = sht._lib.launch_count() - n0
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3, wall], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec, wall = float(t[0].item()), float(t[1].item())
    # per-stage device time of one more simulation (CUDA events between the pieces the library runs in sequence)
    # (the pieces run one after the other on the calling stream here; the first pass is untimed: in the pipeline above the
    # polarization solve lives on its own stream, whose cached allocations this stream cannot reuse)
    for idx in (1000 * rank + 900, 1000 * rank + 901):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        tmap = lib['sims'].get_sim_tmap_dev(idx)
        qmap, umap = lib['sims'].get_sim_pmap_dev(idx)
        ev[1].record()
        tlm = lib['cinv_t'].apply_ivf_dev(tmap)
        ev[2].record()
        elm, blm = lib['cinv_p'].apply_ivf_dev([qmap, umap])
        ev[3].record()
        cl_d = {k: sht.dev_fl(lib['ivfs_raw'].cl[k], lmax) for k in ('tt', 'ee', 'bb', 'te')}
        qe = q._engine(lmax)
        twf = sht.alm_combine([(tlm, cl_d['tt']), (elm, cl_d['te'])])
        ewf = sht.alm_combine([(elm, cl_d['ee']), (tlm, cl_d['te'])])
        G, C = qe.p(tlm, elm, blm, twf, ewf, sht.almxfl(blm, cl_d['bb']))
        ev[4].record()
        torch.cuda.synchronize()
    st = [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
    it_t1, it_p1 = int(lib['cinv_t'].chain.niter), int(lib['cinv_p'].chain.niter)
    assert bool(torch.isfinite(G).all()) and float(torch.linalg.norm(G)) > 0
    f_sim = F0(lmax, NSIDE) + Fs(lmax, NSIDE)                            # synthesis of the simulated T and (Q, U) maps
    # algorithmic: SURVEY.md section 8d ('p' = F0 + 6 Fs: the reference analyses the T and P products separately);
    # executed: what runs here (one merged analysis, three preconditioner applications per multigrid stage)
    ft, fp = cg_flops(lmax, float(np.mean(its_t)), float(np.mean(its_p)), algorithmic=True)
    flop = ft + fp + (F0(lmax, NSIDE) + 6 * Fs(lmax, NSIDE)) + f_sim
    fte, fpe = cg_flops(lmax, float(np.mean(its_t)), float(np.mean(its_p)), algorithmic=False)
    flop_exec = fte + fpe + (F0(lmax, NSIDE) + 5 * Fs(lmax, NSIDE)) + f_sim
    tfl = flop * nsims / sec / 1e12                                       # per GPU: every rank does nsims in `sec`
    tfl_exec = flop_exec * nsims / sec / 1e12
    res = {"nside": NSIDE, "lmax_ivf": lmax, "lmax_qlm": lmax, "sims_per_rank": nsims,
           "sims_per_s": world * nsims / sec, "sims_per_s_wall": world * nsims / wall, "ms_per_sim_per_gpu": 1e3 * sec / nsims,
           "cg_iterations": {"T": its_t, "P": its_p}, "eps_min": 1e-5,
           "final_eps": {"T": float(lib['cinv_t'].chain.last_monitor.trace[-1][1]), "P": float(lib['cinv_p'].chain.last_monitor.trace[-1][1])},
           "masked_cg_iter_per_s": {"T": world * it_t1 / (st[1] * 1e-3), "P": world * it_p1 / (st[2] * 1e-3),
                                    "note": "top-level iterations of one solve / its device time (calc_prep and apply_fini "
                                            "included), x N ranks each filtering its own simulation"},
           "stage_ms_one_sim": {"simulate_TQU_maps": st[0], "cinv_t": st[1], "cinv_p": st[2], "qe_p": st[3],
                                "cg_iterations": {"T": it_t1, "P": it_p1}},
           "algorithmic_flop_per_sim": flop, "algorithmic_tflops_per_gpu": tfl, "frac_of_fp64_nominal_full_volume": tfl / peak_nominal,
           "executed_transforms_flop_per_sim": flop_exec, "executed_transforms_tflops_per_gpu": tfl_exec,
           "frac_of_fp64_nominal_executed_transforms": tfl_exec / peak_nominal,
           "flop_note": "algorithmic = SURVEY.md section 8d counts (full (l, m, ring pair) volume of every transform of the "
                        "reference's algorithm); executed_transforms = the transforms this implementation runs (merged QE "
                        "analysis, 3 instead of 4 preconditioner applications per multigrid stage), still full volume each",
           "kernel_launches_per_sim": launches / max(nsims, 1), "setup_s": t_setup,
           "pipeline": "maps.cmb_maps_nlev (Philox-drawn CMB + noise, synthesised on the GPU) -> filt_cinv.cinv_t / cinv_p "
                       "(reference default chains) -> library_cinv_sepTP -> library_ftl (lmin %d) -> qest.library_sepTP 'p'; "
                       "filtered alms cached to disk asynchronously (PLK_CACHE_FORMAT=%s); mean-field sum reduced over "
                       "ranks inside the timed region" % (LMIN_IVF, os.environ.get('PLK_CACHE_FORMAT', 'fits'))}
    if with_dist_cg and world > 1:
        try:
            res["dist_cg"] = run_dist_cg(lib, lmax, rank, world, dist)
        except Exception as ex:
            import traceback
            traceback.print_exc(file=sys.stderr)
            res["dist_cg"] = {"failed": repr(ex)}
    del lib
    # the alm caches of this lmax (3 x 34-72 MB per simulation and rank) are not needed any more
    barrier()
    if rank == 0:
        import shutil
        shutil.rmtree(os.path.join(tmp, 'lmax%d' % lmax), ignore_errors=True)
    return res


# ------------------------------------------------------------------------------------------------ config 5 (N > 1)
def run_dist(rank, world, dist, steps=3, warmup=2):
    """One 'p' estimate at nside 4096 / lmax_ivf 4000 / lmax_qlm 5000 with every transform m-partitioned over the N
    GPUs; rank 0 also runs the single-GPU plan on the same inputs: bit-identity and speed-up."""
    import torch
    from plancklens_b200 import dist_sht, qest, sht
    nside, lmax, lmax_qlm = DIST_CFG
    cls, transf, ftl, fel, fbl = fiducial(lmax, beam_amin=1.4, nlev_t=5., nlev_p=5. * np.sqrt(2.))
    bars = filtered_sim(0, lmax, cls, transf, (ftl, fel, fbl), nlev_t=5., nlev_p=5. * np.sqrt(2.))
    tbar, ebar, bbar = [sht.dev_alm(x) for x in bars]
    cl_d = {k: sht.dev_fl(cls[k], lmax) for k in ('tt', 'ee', 'bb', 'te')}
    twf = sht.alm_combine([(tbar, cl_d['tt']), (ebar, cl_d['te'])])
    ewf = sht.alm_combine([(ebar, cl_d['ee']), (tbar, cl_d['te'])])
    bwf = sht.almxfl(bbar, cl_d['bb'])
    qe = qest.qe_device(nside, lmax, lmax_qlm, plan_ivf=dist_sht.DistPlan(nside, lmax), plan_qlm=dist_sht.DistPlan(nside, lmax_qlm))

    def sync():
        dist.barrier()
        torch.cuda.synchronize()
    for _ in range(warmup):
        G, C = qe.p(tbar, ebar, bbar, twf, ewf, bwf)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        G, C = qe.p(tbar, ebar, bbar, twf, ewf, bwf)
    e1.record()
    sync()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    qe.plan_ivf.enable_timing(); qe.plan_qlm.enable_timing()
    qe.p(tbar, ebar, bbar, twf, ewf, bwf)
    mine = {'rank': rank, 'ivf': {k: round(v, 2) for k, v in qe.plan_ivf.stage_times().items()},
            'qlm': {k: round(v, 2) for k, v in qe.plan_qlm.stage_times().items()}}
    qe.plan_ivf.enable_timing(False); qe.plan_qlm.enable_timing(False)
    stages = [None] * world
    dist.all_gather_object(stages, mine)
    out = {"nside": nside, "lmax_ivf": lmax, "lmax_qlm": lmax_qlm, "n_gpus": world, "ms_per_estimate": ms, "steps": steps}
    sync()
    if rank == 0:
        # the single-GPU plan on the same inputs (other ranks idle): reference point and bit-identity check.  Same
        # kernels on both sides: the m-partitioned ring stage multiplies the legs in separate kernels, so the single-GPU
        # estimate is evaluated that way too (its default folds the products into the analysis ring kernel)
        os.environ['PLK_QE_FUSED'] = '0'
        ref = qest.qe_device(nside, lmax, lmax_qlm)
        for _ in range(2):
            Gr, Cr = ref.p(tbar, ebar, bbar, twf, ewf, bwf)
        torch.cuda.synchronize()
        e0.record()
        Gr, Cr = ref.p(tbar, ebar, bbar, twf, ewf, bwf)
        e1.record()
        torch.cuda.synchronize()
        ms1 = e0.elapsed_time(e1)
        os.environ.pop('PLK_QE_FUSED', None)
        flop = F0(lmax, nside) + 4 * Fs(lmax, nside) + Fs(lmax_qlm, nside)
        out.update({"ms_single_gpu": ms1, "speedup": ms1 / ms, "efficiency": ms1 / ms / world,
                    "bit_identical_to_single_gpu": bool(torch.equal(G, Gr) and torch.equal(C, Cr)),
                    "rel_l2_vs_single_gpu": float(torch.linalg.norm(G - Gr) / torch.linalg.norm(Gr)),
                    "algorithmic_tflops_total": flop / (ms * 1e-3) / 1e12, "stage_ms_one_estimate": stages,
                    "exchange": "peer stores over NVLink fused into legendre_synth / ring_anal kernels; qlm rows summed "
                                "with one all-reduce"})
    dist.barrier()
    del qe
    sht.clear_plans()
    return out


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=12)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', type=str, default='b200')
    ap.add_argument('--cpu-mstep', type=int, default=16)
    ap.add_argument('--ref-budget-s', type=float, default=150.0, help="wall-clock budget of the --impl reference arm")
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-target', action='store_true', help="skip the masked-sky CG-filtered pipeline (extra.target)")
    ap.add_argument('--target-sims', type=int, default=6, help="timed simulations per rank and lmax of extra.target")
    ap.add_argument('--no-dist', action='store_true', help="skip the m-partitioned nside-4096 estimate (extra.dist, N > 1)")
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    os.environ.setdefault('PLK_CACHE_FORMAT', 'npy')     # .npy caches under the reference's file names (FITS: +0.15 s per alm)
    from plancklens_b200 import qest, sht

    cls, transf, ftl, fel, fbl = fiducial(LMAX_IVF)
    # pool of distinct simulated skies per rank, pinned on the host and resident on the device
    host_sims, dev_sims = [], []
    for i in range(NPOOL):
        alms = filtered_sim(rank * 1000 + i, LMAX_IVF, cls, transf, (ftl, fel, fbl))
        pins = [torch.from_numpy(a).pin_memory() for a in alms]
        host_sims.append([p.numpy() for p in pins])
        dev_sims.append([p.cuda() for p in pins])
    ivfs = mem_ivfs(host_sims, cls, NSIDE)
    import tempfile
    # one lib_dir for all ranks (rank 0 writes the hash files, the others wait at the library's barrier)
    tmp = os.path.join(tempfile.gettempdir(), 'plk_bench_%s_%s' % (os.environ.get('MASTER_PORT', 'single'), os.getppid()))
    lib = qest.library_sepTP(os.path.join(tmp, 'qlms_dd'), ivfs, ivfs, cls['te'], NSIDE, lmax_qlm=LMAX_QLM)
    f2 = lib.f2map2
    qe = lib._engine(LMAX_IVF)
    nalm_q = alm_size(LMAX_QLM)
    mf = [torch.zeros(nalm_q, dtype=torch.complex128, device='cuda') for _ in range(2)]

    def step_device(i):
        dt, de, db = dev_sims[i % NPOOL]
        twf, ewf, bwf = f2.wf_device(i, 'p', (dt, de, db))
        G, C = qe.p(dt, de, db, twf, ewf, bwf)
        sht.alm_axpy(mf[0], G, 1.0)
        sht.alm_axpy(mf[1], C, 1.0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput
    for i in range(args.warmup):
        step_device(i)
    barrier()
    sht.profile_enable(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = sht._lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_device(args.warmup + i)
    if world > 1:
        for v in mf:                                   # mean-field sum over ranks (qest.py:239-243)
            dist.reduce(torch.view_as_real(v), dst=0)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = sht._lib.launch_count() - n0
    prof = sht.profile_read()
    sht.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    t_dev = torch.tensor([ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    ms_dev = float(t_dev.item())

    # ---------------- end to end through the public API (numpy in pinned host memory -> numpy out)
    for _ in lib.eval_qlms('p', range(3)):
        pass
    barrier()
    t0 = time.perf_counter()
    for _, G, C in lib.eval_qlms('p', range(args.warmup, args.warmup + args.steps)):
        pass
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    s_e2e = float(t_e2e.item())
    assert np.all(np.isfinite(G[:100])) and np.any(G != 0)

    sm_max = (clocks or {}).get('sm_max_mhz') or 1965.0
    sm_run = (clocks or {}).get('sm_mhz') or sm_max
    peak_nominal = 148 * 64 * 2 * sm_max * 1e6 / 1e12          # FP64 FMA pipe: 64 FMA / clk / SM
    peaks = measured_peaks()

    # ---------------- the other estimators of the metric ('ptt', 'p_p') and the per-kernel rooflines, single GPU
    extra = {}
    kernels = {}
    if world == 1:
        for key in ('ptt', 'p_p'):
            def step_k(i, key=key):
                dt, de, db = dev_sims[i % NPOOL]
                c = f2._cl_dev()
                if key == 'ptt':
                    return qe.ptt(dt, sht.almxfl(dt, c['tt']))
                return qe.p_p(de, db, sht.almxfl(de, c['ee']), sht.almxfl(db, c['bb']))
            for i in range(2):
                step_k(i)
            torch.cuda.synchronize()
            nk = max(4, args.steps // 2)
            e0.record()
            for i in range(nk):
                step_k(i)
            e1.record()
            torch.cuda.synchronize()
            extra['%s_qlm_per_s' % key] = nk / (e0.elapsed_time(e1) * 1e-3)
        plan = sht.get_plan(NSIDE, LMAX_IVF)
        # flop per launch: SURVEY.md section 8d (8 / 24 per unit); the gradient-only kernel executes 16 per unit
        fl_k = {'synth_spin0': (F0(LMAX_IVF, NSIDE), 0), 'anal_spin0': (F0(LMAX_IVF, NSIDE), 0),
                'synth_spins': (Fs(LMAX_IVF, NSIDE), 2), 'anal_spins': (Fs(LMAX_IVF, NSIDE), 1),
                'synth_grad': (Fs(LMAX_IVF, NSIDE) * 16. / 24., 1)}
        for k, v in prof.items():
            if v[0] == 0:
                continue
            msl = v[1] / v[0]
            full = fl_k[k][0] / (msl * 1e-3) / 1e12
            act = plan.active_fraction(fl_k[k][1])
            kernels['legendre_' + k] = {"bound": "fp64", "launches": v[0], "ms_per_launch": msl, "tflops_full_volume": full,
                                        "executed_share_of_volume": act, "achieved": full * act, "peak": peak_nominal,
                                        "frac": full * act / peak_nominal, "unit": "TFLOP/s"}
        # ring-FFT stage on its own: HBM roofline (phase array in + map out, or the reverse), L2 flushed by the sizes
        X = plan.new_phase()
        m = torch.empty(plan.npix, dtype=torch.float64, device='cuda')
        hbm = peaks.get('hbm_gbs')
        for name, fn in (('ring_synth', lambda: plan.ring_synth(X, out=m)), ('ring_anal', lambda: plan.ring_anal(m, X=X))):
            fn(); torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            msl = e0.elapsed_time(e1) / 5
            nbytes = plan.nring * (LMAX_IVF + 1) * 16 + plan.npix * 8
            kernels[name] = {"bound": "hbm", "ms_per_component": msl, "algorithmic_bytes": nbytes, "achieved": nbytes / (msl * 1e-3) / 1e9,
                             "peak": hbm, "frac": (nbytes / (msl * 1e-3) / 1e9 / hbm) if hbm else None, "unit": "GB/s",
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs" if hbm else "MEASURED_PEAKS.json absent"}
        del X, m
    extra['kernels'] = kernels

    # ---------------- north_star target: masked-sky CG-filtered 'p' QE, simulations sharded over ranks
    lib = qe = f2 = None
    dev_sims_keep = dev_sims
    if not args.no_target:
        try:
            import contextlib
            with contextlib.redirect_stdout(sys.stderr):
                mask, z = synthetic_sky_model(NSIDE)
                extra['target'] = {'fsky': float(mask.mean())}
                for lm in TARGET_LMAX:
                    extra['target']['lmax%d' % lm] = run_target(lm, args.target_sims, tmp, mask, z, rank, world, dist, peak_nominal,
                                                                 with_dist_cg=(lm == TARGET_LMAX[0] and not args.no_dist))
                    sht.clear_plans()
                    import gc
                    gc.collect()
                    torch.cuda.empty_cache()
        except Exception as ex:
            import traceback
            traceback.print_exc(file=sys.stderr)
            extra.setdefault('target', {})['failed'] = repr(ex)
    # ---------------- config 5: one estimate m-partitioned over the GPUs of the box
    if world > 1 and not args.no_dist:
        try:
            import contextlib
            with contextlib.redirect_stdout(sys.stderr):
                extra['dist'] = run_dist(rank, world, dist)
        except Exception as ex:
            import traceback
            traceback.print_exc(file=sys.stderr)
            extra['dist'] = {'failed': repr(ex)}

    if rank == 0:
        fs = Fs(LMAX_IVF, NSIDE)
        cnt, tot = prof['synth_spins']
        k_ms = tot / max(cnt, 1)
        full = fs / (k_ms * 1e-3) / 1e12
        share = {k: round(v[1] / ms_dev, 4) for k, v in prof.items()}
        act = sht.get_plan(NSIDE, LMAX_IVF).active_fraction(2)
        traffic, traffic_src = ncu_traffic('legendre_synth_kernel<1, 4, 0>')
        line = {
            "metric": "QE qlms/sec ('p', nside 2048, lmax 2048)", "value": world * args.steps / (ms_dev * 1e-3), "unit": "qlm/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(),
            "e2e": {"value": world * args.steps / s_e2e, "unit": "qlm/s",
                    "h2d_bytes_per_step": 3 * alm_size(LMAX_IVF) * 16, "d2h_bytes_per_step": 2 * nalm_q * 16,
                    "api": "plancklens_b200.qest.library.eval_qlms('p', idxs): numpy alms in pinned host memory in, numpy qlm out; "
                           "H2D of sim i+1 / transforms of sim i / D2H of sim i-1 overlap on two streams, every copy inside the timed region"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "kernel": "legendre_synth_kernel<spin,NR=4>", "achieved": full * act, "peak": peak_nominal,
                         "unit": "TFLOP/s", "frac": full * act / peak_nominal,
                         "peak_source": "nominal FP64 FMA pipe: 148 SM x 64 FMA/clk x 2 flop x %.0f MHz (max SM clock sampled by "
                                        "nvidia-smi during the timed region; median under load %.0f MHz).  MEASURED_PEAKS.json holds "
                                        "HBM and bf16 peaks only, neither bounds this kernel" % (sm_max, sm_run),
                         "executed_share_of_volume": act, "achieved_full_volume": full, "frac_full_volume": full / peak_nominal,
                         "note": "achieved = DFMA rate the kernel sustains: 24 flop x the (l, m, ring pair) volume it walks / launch "
                                 "time; it skips the share of the volume below the 2^-60 start threshold near the poles.  "
                                 "achieved_full_volume credits the whole volume (the SURVEY.md section 8d count)",
                         "launch_ms": k_ms, "launches_timed": cnt, "flop_per_launch_full_volume": fs,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": 2 * alm_size(LMAX_IVF) * 16 + 2 * (4 * NSIDE - 1) * (LMAX_IVF + 1) * 16,
                         "kernel_share_of_step": share},
            "clocks": clocks,
            "extra": extra,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle import ref_sht
                ref_sht.build()
                t_cpu, desc = cpu_port_step([a for a in host_sims[0]], cls, args.cpu_mstep)
                line["cpu_baseline"] = {"value": 1.0 / t_cpu, "unit": "qlm/s", "cores": ref_sht.max_threads(), "kind": "port",
                                        "sample": desc}
            except Exception as ex:   # the baseline is informative only
                line["cpu_baseline"] = {"value": None, "unit": "qlm/s", "cores": None, "kind": "port", "sample": "failed: %r" % (ex,)}
        print(json.dumps(line))
    del dev_sims_keep
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
