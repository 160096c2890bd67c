#!/usr/bin/env python
"""Turn the CSV log of `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file X`
into the per-kernel launch list kept under profiles/ (shares of the step, cold caches, serialised launches).

    python profiles/ncu_launches.py gpurun_out/launches.csv "<command line that was profiled>" > profiles/r01_launches.txt"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
iid = hdr.index("ID")
per = OrderedDict()
for r in rows[1:]:
    d = per.setdefault(r[iid], {"name": r[ik]})
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    if r[im] == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    else:
        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
    d[r[im]] = v
agg = OrderedDict()
for d in per.values():
    a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0)
    a[3] += d.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
print("# %s" % (sys.argv[2] if len(sys.argv) > 2 else "ncu launch list"))
print("# every kernel launch of the run, cold caches, serialised: compare SHARES, not absolutes")
print("# %d launches, %.1f ms of kernel time" % (len(per), tot / 1000.0))
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-70s n=%4d %10.1f us total (%4.1f%%)  avg %8.1f us  dram rd %7.1f MB  wr %7.1f MB per launch" %
          (name[:70], a[0], a[1], 100.0 * a[1] / tot, a[1] / a[0], a[2] / a[0], a[3] / a[0]))
