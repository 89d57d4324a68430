#!/usr/bin/env python
"""Per-kernel launch statistics of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
   python tools/launch_times.py gpurun_out/launches.csv [...]"""
import collections
import csv
import statistics
import sys

for fn in sys.argv[1:]:
    rows = [r for r in csv.reader(open(fn)) if len(r) > 14 and r[0].isdigit()]
    print(fn, len(rows), "launches")
    agg = collections.OrderedDict()
    for r in rows:
        agg.setdefault(r[4].split("(")[0].replace("<unnamed>::", "").replace("void ", "")[:60] + " grid " + r[8], []).append(float(r[-1]))
    for k, v in agg.items():
        print("  %-90s n=%4d  median %9.1f us  min %9.1f  total %9.3f ms" % (k, len(v), statistics.median(v) / 1e3, min(v) / 1e3, sum(v) / 1e6))
