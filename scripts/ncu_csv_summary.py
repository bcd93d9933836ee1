"""Prints one line per kernel launch from an `ncu --csv` log (durations in us, DRAM bytes in GB)."""
import csv, sys
from collections import OrderedDict
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
d = OrderedDict()
for row in csv.DictReader(lines):
    d.setdefault((row['ID'], row['Kernel Name'][:48]), {})[row['Metric Name']] = (row['Metric Value'], row['Metric Unit'])
def num(m, k):
    v = m.get(k)
    return float(v[0].replace(',', '')) if v else float('nan')
for (i, k), m in d.items():
    t, u = m['gpu__time_duration.sum']
    t = float(t.replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(u, 1e-3)
    print(f"{i:>3} {k:<48} {t:10.1f} us  rd {num(m,'dram__bytes_read.sum')/1e9:7.2f}  wr {num(m,'dram__bytes_write.sum')/1e9:7.2f}  fmaheavy {num(m,'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed'):5.1f}%  warps {num(m,'sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f}%")
