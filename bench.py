#!/usr/bin/env python
"""Benchmark of the plancklens hot path on B200: quadratic-estimator evaluations per second.

Contract (driver): `python bench.py --gpus N --steps K --warmup W` (under torchrun for N > 1) prints ONE JSON line.

Workload (BASELINE.json configs[1]): idealized full-sky MV 'p' lensing QE, nside 2048, lmax_ivf = lmax_qlm = 2048,
synthetic Gaussian CMB + white-noise skies filtered isotropically (params/idealized_example.py of the reference
with the FFP10 fiducial spectra, 5' beam, 35 / 55 uK-arcmin).  One step = `qest.library.get_sim_qlm('p', idx)` for
one simulation starting from its cached inverse-variance filtered alms (what `run_qlms.py -k p -dd` does after
`-ivt -ivp`): 1 spin-0 + 4 spin-s syntheses, per-pixel products, spin-1 analysis -> (glm, clm).
  value : steps/s with the filtered alms resident in HBM (CUDA events, max over ranks)
  e2e   : the same through `qest.library.eval_qlms` (the pipelined form of `eval_qlm`) with numpy (pinned host) inputs
          and numpy outputs, every H2D + D2H inside the timed region
  roofline : the dominant kernel, `legendre_synth_kernel<spin>`; achieved = 24 flop x N_lm x 2 nside per launch
           / its mean CUDA-event duration inside the timed region; bound = FP64 FMA pipe
  cpu_baseline : the CPU oracle port of the same step on a bounded sample (every MSTEP-th m), host cores stated
`--impl reference` prints the same line for the CPU port alone (the reference itself needs healpy, absent here).
Multi-GPU: simulations are sharded over ranks (idx % N == rank, reference: examples/run_qlms.py:72), weak scaling,
one NCCL reduce of the accumulated qlm (the mean-field sum of qest.py:239-243) inside the timed region.
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


def alm_size(lmax):
    return (lmax + 1) * (lmax + 2) // 2


def n_lm(lmax):
    return sum(lmax - m + 1 for m in range(lmax + 1))


def fiducial(lmax):
    from plancklens_b200 import hp, utils
    cls = utils.camb_clfile(os.path.join(ROOT, 'plancklens_b200', 'data', 'cls', 'FFP10_wdipole_lensedCls.dat'), lmax=lmax)
    transf = hp.gauss_beam(BEAM_AMIN / 60. / 180. * np.pi, lmax=lmax)
    ftl = utils.cli(cls['tt'] + (NLEV_T / 60. / 180. * np.pi / transf) ** 2)
    fel = utils.cli(cls['ee'] + (NLEV_P / 60. / 180. * np.pi / transf) ** 2)
    fbl = utils.cli(cls['bb'] + (NLEV_P / 60. / 180. * np.pi / transf) ** 2)
    for f in (ftl, fel, fbl):
        f[:LMIN_IVF] = 0.
    return cls, transf, ftl, fel, fbl


def filtered_sim(idx, lmax, cls, transf, fls):
    """Inverse-variance filtered alms of one synthetic sky: f_l (a_lm + n_lm / b_l), correlated T/E draw
    (reference recipe: sims/phas.py:162-168, sims/cmbs.py:35-69, white noise of sims/maps.py:136-173 in harmonic space)."""
    from plancklens_b200 import hp
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
    for alm, nlev, fl in ((tlm, NLEV_T, fls[0]), (elm, NLEV_P, fls[1]), (blm, NLEV_P, fls[2])):
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


# ------------------------------------------------------------------------------------------------ CPU port
def cpu_port_step(sims, cls, mstep):
    """One 'p' estimate with the CPU oracle on a bounded sample: Legendre stages on every mstep-th m (scaled by the
    sampled share of the (l, m) work), ring FFTs and pixel products on one component each (scaled by the count).
    Returns (estimated seconds for the full step, description)."""
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
    t0 = time.time()
    X0, _ = ref_sht.legendre_synth(NSIDE, 0, lmax, lmax, tbar, mstep=mstep)
    ref_sht.legendre_synth(NSIDE, 1, lmax, lmax, almxfl(twf, -np.sqrt(l * (l + 1))), np.zeros_like(twf), mstep=mstep)
    ref_sht.legendre_synth(NSIDE, 2, lmax, lmax, 0.5 * ebar, 0.5 * bbar, mstep=mstep)
    f3 = np.sqrt(np.maximum((l - 2) * (l + 3), 0)); f3[:3] = 0
    f1 = np.sqrt(np.maximum((l + 2) * (l - 1), 0)); f1[:1] = 0
    ref_sht.legendre_synth(NSIDE, 3, lmax, lmax, almxfl(ewf, f3), almxfl(bwf, f3), mstep=mstep)
    X1, X2 = ref_sht.legendre_synth(NSIDE, 1, lmax, lmax, almxfl(ewf, f1), almxfl(bwf, f1), mstep=mstep)
    ref_sht.legendre_anal(NSIDE, 1, LMAX_QLM, LMAX_QLM, X1, X2, mstep=mstep)     # the reference runs two analyses
    ref_sht.legendre_anal(NSIDE, 1, LMAX_QLM, LMAX_QLM, X1, X2, mstep=mstep)
    t_leg = (time.time() - t0) * scale
    t0 = time.time()
    m = ref_sht.phase2map(NSIDE, X0)
    t_s = time.time() - t0
    t0 = time.time()
    ref_sht.map2phase(NSIDE, m, LMAX_QLM)
    t_a = time.time() - t0
    t0 = time.time()
    g = m * m; c = m * m + g; c -= (m + 1j * m).real * g     # stand-in for the ~8 full-map numpy passes of qest.py:256-278
    t_pix = (time.time() - t0) * 3
    total = t_leg + 9 * t_s + 4 * t_a + t_pix
    return total, "Legendre on every %d-th m (x%.1f), ring FFT 1 of 9+4 components, pixel passes x3" % (mstep, scale)


def run_reference(args):
    """--impl reference: the CPU port of the step (healpy is not installable here), rank 0 only."""
    if int(os.environ.get('RANK', 0)) != 0:
        return
    if int(os.environ.get('WORLD_SIZE', 1)) > 1:
        # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm runs on rank 0 alone and gets all host cores
        os.environ['OMP_NUM_THREADS'] = str(os.cpu_count())
    from oracle import ref_sht
    ref_sht.build()
    cls, transf, ftl, fel, fbl = fiducial(LMAX_IVF)
    sims = filtered_sim(0, LMAX_IVF, cls, transf, (ftl, fel, fbl))
    cores = ref_sht.max_threads()
    ts = []
    desc = ''
    for i in range(args.warmup + args.steps):
        t, desc = cpu_port_step(sims, cls, args.cpu_mstep)
        if i >= args.warmup:
            ts.append(t)
    t = float(np.median(ts))
    line = {"impl": "reference", "metric": "QE qlms/sec ('p', nside 2048, lmax 2048)", "value": 1.0 / t, "unit": "qlm/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(),
            "cpu_baseline": {"value": 1.0 / t, "unit": "qlm/s", "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": 1.0 / t, "unit": "qlm/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU oracle port (oracle/csht.c + numpy FFT): the reference's own path needs healpy, which is not installed and not installable offline"}
    print(json.dumps(line))


def workload_config():
    return {"workload": "idealized full-sky MV 'p' lensing QE from cached inverse-variance filtered alms "
                        "(BASELINE.json configs[1]): nside 2048, lmax_ivf 2048, lmax_qlm 2048, 1 spin-0 + 4 spin-s "
                        "syntheses + spin-1 analysis per estimate",
            "nside": NSIDE, "lmax_ivf": LMAX_IVF, "lmax_qlm": LMAX_QLM, "key": "p",
            "l2_policy": "every step streams ~6 GB of maps and phase arrays (>> 126 MB L2); inputs rotate over %d sims" % NPOOL,
            "parallelism": "simulations sharded over ranks (idx % N == rank)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=12)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', type=str, default='b200')
    ap.add_argument('--cpu-mstep', type=int, default=32)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-cg', action='store_true', help="skip the masked-sky CG part of the metric (N = 1 only, ~2 min)")
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

    # ---------------- the other estimators of the metric ('ptt', 'p_p'), device resident, rank 0 of a single-GPU run
    extra = {}
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
        F0 = 8.0 * n_lm(LMAX_IVF) * 2 * NSIDE
        Fs_ = 24.0 * n_lm(LMAX_IVF) * 2 * NSIDE
        # flop per launch: SURVEY.md section 8d (8 / 24 per unit); the gradient-only kernel executes 16 per unit and is
        # credited with what it executes
        fl_k = {'synth_spin0': F0, 'anal_spin0': F0, 'synth_spins': Fs_, 'anal_spins': Fs_, 'synth_grad': Fs_ * 16. / 24.}
        extra['legendre_kernels'] = {k: {"launches": v[0], "ms_per_launch": v[1] / max(v[0], 1),
                                         "tflops": fl_k[k] / (v[1] / max(v[0], 1) * 1e-3) / 1e12}
                                     for k, v in prof.items() if v[0] > 0}
        if not args.no_cg:
            import contextlib
            sys.path.insert(0, os.path.join(ROOT, 'scripts'))
            try:
                import bench_cg
                with contextlib.redirect_stdout(sys.stderr):
                    extra['masked_cg'] = bench_cg.run(NSIDE, LMAX_IVF, True, True, verbose=False)
            except Exception as ex:
                extra['masked_cg'] = {'failed': repr(ex)}

    if rank == 0:
        Fs = 24.0 * n_lm(LMAX_IVF) * 2 * NSIDE
        cnt, tot = prof['synth_spins']
        k_ms = tot / max(cnt, 1)
        peak_meas = sht.fp64_peak_tflops(3)
        sm_max = (clocks or {}).get('sm_max_mhz') or 1965.0
        peak_nominal = 148 * 64 * 2 * sm_max * 1e6 / 1e12
        achieved = Fs / (k_ms * 1e-3) / 1e12
        share = {k: round(v[1] / ms_dev, 4) for k, v in prof.items()}
        act = sht.get_plan(NSIDE, LMAX_IVF).active_fraction(2)
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
            "roofline": {"bound": "fp64", "kernel": "legendre_synth_kernel<spin,NR=4>", "achieved": achieved, "peak": peak_meas,
                         "unit": "TFLOP/s", "frac": achieved / peak_meas,
                         "peak_source": "DFMA microbenchmark in libplk_b200 (plk_fp64_peak) run in this process; MEASURED_PEAKS.json "
                                        "holds only HBM and bf16 peaks, neither bounds this kernel",
                         "peak_nominal": peak_nominal, "frac_of_nominal": achieved / peak_nominal,
                         "executed_share_of_volume": act, "achieved_executed": achieved * act,
                         "frac_executed": achieved * act / peak_meas,
                         "note": "achieved counts the full (l, m, ring-pair) volume of SURVEY.md section 8d; the kernel skips the "
                                 "share below the 2^-120 start threshold near the poles (executed_share_of_volume), so frac can "
                                 "exceed 1 -- frac_executed is the DFMA rate actually sustained",
                         "launch_ms": k_ms, "launches_timed": cnt, "flop_per_launch": Fs,
                         "traffic": 851e6, "traffic_source": "ncu dram__bytes_read+write per launch, profiles/r01_ncu_summary.md",
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
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
