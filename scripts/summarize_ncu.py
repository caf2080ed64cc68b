"""Turns ncu outputs brought back from the GPU box into the tracked markdown summary under profiles/."""
import csv, subprocess, sys, collections, io, re

launch_csv, full_rep, out_md, tag = sys.argv[1:5]
rows = [r for r in csv.reader(open(launch_csv)) if r]
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]
ik, iv, im = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= iv or r[im] != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', r[ik]).replace('void ', '').replace('plk::', '')
    unit = r[hdr.index('Metric Unit')]
    v = float(r[iv].replace(',', ''))
    v_us = v / 1e3 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1e3)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v_us
tot = sum(a[1] for a in agg.values())
lines = ['# ncu summary %s' % tag, '',
         'Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline`',
         '(cold-cache, serialised launches: compare SHARES, not absolutes).', '',
         '| kernel | launches | total us | share |', '|---|---:|---:|---:|']
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append('| `%s` | %d | %.0f | %.1f %% |' % (k, n, t, 100 * t / tot))
lines += ['', '## `ncu --set full` per-launch metrics (same command, `-k regex:legendre_|ring_`)', '']
raw = subprocess.run(['ncu', '-i', full_rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h, u = rr[0], rr[1]
ix = {n: i for i, n in enumerate(h)}
want = [('gpu__time_duration.sum', 'time'), ('launch__grid_size', 'grid'), ('launch__registers_per_thread', 'regs'),
        ('launch__shared_mem_per_block_dynamic', 'dyn smem'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
        ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'fp64 pipe active %'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'),
        ('dram__bytes_read.sum', 'dram read'), ('dram__bytes_write.sum', 'dram write'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram % of peak'),
        ('sm__cycles_elapsed.avg', 'sm cycles')]
lines.append('| kernel | ' + ' | '.join(w[1] for w in want) + ' |')
lines.append('|---|' + '---:|' * len(want))
seen = collections.Counter()
for r in rr[2:]:
    name = re.sub(r'\(.*', '', r[ix['Kernel Name']]).replace('void ', '').replace('plk::', '')
    seen[name] += 1
    if seen[name] > 2:
        continue
    vals = []
    for m, _ in want:
        vals.append('%s %s' % (r[ix[m]], u[ix[m]]) if m in ix else '-')
    lines.append('| `%s` | ' % name + ' | '.join(vals) + ' |')
open(out_md, 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines))
