"""Throughput of the north_star target pipeline (bench.py build_target: masked-sky cinv_t + cinv_p -> 'p' estimate) for
the scheduling in force -- PLK_TP_CONCURRENT (T and P filters side by side, default 1), PLK_CG_PRIO (preconditioner
graphs on a high-priority stream, default 1) -- with a digest of one estimate to show the result does not depend on it.

  python scripts/time_tp.py [lmax] [nsims]"""
import contextlib
import hashlib
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

lmax = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
nsims = int(sys.argv[2]) if len(sys.argv) > 2 else 4
os.environ.setdefault('PLK_CACHE_FORMAT', 'npy')
tmp = tempfile.mkdtemp(prefix='plk_tp_')
with contextlib.redirect_stdout(open(os.devnull, 'w')):
    mask, z = bench.synthetic_sky_model(bench.NSIDE)
    lib = bench.build_target(lmax, tmp, mask, z)
    q = lib['qlms_dd']
    for i in range(2):
        q.get_sim_qlm_dev('p', i)
    torch.cuda.synchronize()
    trace = []

    def wrap(name, obj):
        inner = obj.apply_ivf_dev

        def f(*a, **k):
            import time
            st = torch.cuda.current_stream()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(st)
            h0 = time.perf_counter()
            r = inner(*a, **k)
            h1 = time.perf_counter()
            a1.record(st)
            trace.append((name, a0, a1, h0, h1))
            return r
        obj.apply_ivf_dev = f
    wrap('T', lib['cinv_t'])
    wrap('P', lib['cinv_p'])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    its = []
    e0.record()
    depth = int(os.environ.get('PLK_TP_PREFETCH', '2'))
    idxs = [20 + i for i in range(nsims)]
    digs = []
    for i, idx in enumerate(idxs):
        G, C = q.get_sim_qlm_dev('p', idx, prefetch=idxs[i + 1:i + 1 + depth])
        it = lib['ivfs_raw'].cg_iterations[idx]
        its.append((it['T'], it['P']))
        digs.append(G.clone())
    e1.record()
    torch.cuda.synchronize()
    lib['ivfs'].flush()
for name, a0, a1, h0, h1 in trace:
    print('   %s solve: device %.1f ms, host call %.1f ms (host start %.1f ms after the first)' %
          (name, a0.elapsed_time(a1), 1e3 * (h1 - h0), 1e3 * (h0 - trace[0][3])))
sec = e0.elapsed_time(e1) * 1e-3
dig = hashlib.sha1(b''.join(g.cpu().numpy().tobytes() for g in digs)).hexdigest()[:12]
print('lmax %d concurrent=%s prio=%s prefetch=%s : %.3f sims/s (%.1f ms per simulation), CG iterations (T, P) %s, digest of all G %s' %
      (lmax, os.environ.get('PLK_TP_CONCURRENT', '1'), os.environ.get('PLK_CG_PRIO', '1'), os.environ.get('PLK_TP_PREFETCH', '2'), nsims / sec, 1e3 * sec / nsims, its, dig))
