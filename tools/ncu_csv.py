#!/usr/bin/env python
"""dev tool: condense an `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` log"""
import csv, sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
d = OrderedDict()
for r in rows:
    d.setdefault(r[0], {"k": r[4]})[r[12]] = float(r[14])
for i, v in d.items():
    print(i, v["k"][5:24], "read MB", round(v.get("dram__bytes_read.sum", 0) / 1e6, 1), "write MB",
          round(v.get("dram__bytes_write.sum", 0) / 1e6, 1), "us", round(v.get("gpu__time_duration.sum", 0) / 1e3, 1))
