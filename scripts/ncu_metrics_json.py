"""ncu raw-page CSV (ncu -i x.ncu-rep --page raw --csv) -> profiles/<tag>_ncu_metrics.json, the file bench.py reads
`roofline.traffic` from.  usage: python scripts/ncu_metrics_json.py "<source note>" raw.csv > profiles/r02_ncu_metrics.json"""
import collections
import csv
import json
import re
import sys

note, raw = sys.argv[1], sys.argv[2]
rr = list(csv.reader(open(raw)))
h, u = rr[0], rr[1]
ix = {n: i for i, n in enumerate(h)}


def num(r, m, scale_units=True):
    v = float(r[ix[m]].replace(',', ''))
    unit = u[ix[m]]
    if scale_units:
        v *= {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'us': 1e-3, 'usecond': 1e-3, 'ns': 1e-6, 'nsecond': 1e-6, 'ms': 1.0,
              'msecond': 1.0, 's': 1e3, 'second': 1e3}.get(unit, 1.0)
    return v


out = collections.OrderedDict()
seen = collections.Counter()
for r in rr[2:]:
    name = re.sub(r'\(.*', '', r[ix['Kernel Name']]).replace('void ', '').replace('plk::', '')
    seen[name] += 1
    if seen[name] != 2 and name in out:      # keep the second launch of a kernel (first one is the cold warm-up)
        continue
    out[name] = {'ms': num(r, 'gpu__time_duration.sum'), 'dram_read_bytes': num(r, 'dram__bytes_read.sum'),
                 'dram_write_bytes': num(r, 'dram__bytes_write.sum'),
                 'fp64_pipe_pct': num(r, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', False),
                 'registers': int(num(r, 'launch__registers_per_thread', False))}
json.dump({'source': note, 'kernels': out}, sys.stdout, indent=1)
print()
