"""Turns ncu outputs brought back from the GPU box (gpurun_out/) into the tracked markdown summary under profiles/.

usage: python scripts/summarize_ncu.py <tag> <launches.csv> <raw1.csv> [<raw2.csv> ...] > profiles/<tag>_ncu_summary.md
  launches.csv : ncu --metrics gpu__time_duration.sum --clock-control none ... --csv --log-file launches.csv <cmd>
  rawN.csv     : ncu -i <rep from `ncu --set full --clock-control none`> --page raw --csv
"""
import collections
import csv
import re
import sys

tag, launch_csv, raws = sys.argv[1], sys.argv[2], sys.argv[3:]


def short(name):
    return re.sub(r'\(.*', '', name).replace('void ', '').replace('plk::', '').replace('(bool)', '').replace('(int)', '')


rows = [r for r in csv.reader(open(launch_csv)) if r]
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]
ik, iv, im, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= iv or r[im] != 'gpu__time_duration.sum':
        continue
    v = float(r[iv].replace(',', ''))
    v_us = v / 1e3 if r[iu] in ('ns', 'nsecond') else (v if r[iu] in ('us', 'usecond') else v * 1e3)
    a = agg.setdefault(short(r[ik]), [0, 0.0])
    a[0] += 1
    a[1] += v_us
tot = sum(a[1] for a in agg.values())
print('# ncu summary %s' % tag)
print()
print('## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold-cache and serialised: compare SHARES)')
print()
print('| kernel | launches | total us | share | avg us |')
print('|---|---:|---:|---:|---:|')
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if t / tot < 0.0005:
        continue
    print('| `%s` | %d | %.0f | %.1f %% | %.1f |' % (k, n, t, 100 * t / tot, t / n))
print()
print('## `ncu --set full --clock-control none` per-launch metrics')
print()
want = [('gpu__time_duration.sum', 'time'), ('launch__grid_size', 'grid'), ('launch__registers_per_thread', 'regs'),
        ('launch__shared_mem_per_block_dynamic', 'dyn smem'), ('sm__warps_active.avg.per_cycle_active', 'warps/SM'),
        ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'fp64 pipe %'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue %'),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smem wavefronts %'),
        ('dram__bytes_read.sum', 'dram read'), ('dram__bytes_write.sum', 'dram write'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'),
        ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall wait'),
        ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'stall math'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall long_sb'),
        ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall barrier')]
print('| kernel | ' + ' | '.join(w[1] for w in want) + ' |')
print('|---|' + '---:|' * len(want))
for raw in raws:
    rr = list(csv.reader(open(raw)))
    h, u = rr[0], rr[1]
    ix = {n: i for i, n in enumerate(h)}
    seen = collections.Counter()
    for r in rr[2:]:
        name = short(r[ix['Kernel Name']])
        key = name + r[ix['launch__grid_size']]
        seen[key] += 1
        if seen[key] > 1:
            continue
        vals = []
        for m, _ in want:
            if m not in ix:
                vals.append('-')
                continue
            v = r[ix[m]]
            try:
                v = '%.3g' % float(v.replace(',', ''))
            except ValueError:
                pass
            vals.append('%s %s' % (v, u[ix[m]].replace('/thread', '').replace('/block', '')) if u[ix[m]] not in ('', '%') else v)
        print('| `%s` | ' % name + ' | '.join(vals) + ' |')
