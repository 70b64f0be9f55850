"""Summarise an ncu --csv launch list (gpu__time_duration.sum [+ dram bytes]) per kernel: launches, total and mean time, DRAM GB."""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
per = OrderedDict()
launch = {}
for r in rows[hi + 1:]:
    if len(r) < 15:
        continue
    launch.setdefault(r[0], {"name": r[4].split("(")[0].replace("void ", "").replace("txr::", "")})[r[12]] = float(r[14].replace(",", ""))
for i, m in launch.items():
    d = per.setdefault(m["name"], {"n": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0})
    d["n"] += 1
    d["ns"] += m.get("gpu__time_duration.sum", 0.0)
    d["rd"] += m.get("dram__bytes_read.sum", 0.0)
    d["wr"] += m.get("dram__bytes_write.sum", 0.0)
tot = sum(d["ns"] for d in per.values())
print(f"{'kernel':58s} {'n':>4s} {'total ms':>9s} {'mean us':>9s} {'share':>6s} {'rd GB':>8s} {'wr GB':>8s}")
for k, d in sorted(per.items(), key=lambda kv: -kv[1]["ns"]):
    print(f"{k[:58]:58s} {d['n']:4d} {d['ns'] / 1e6:9.3f} {d['ns'] / d['n'] / 1e3:9.1f} {d['ns'] / tot:6.1%} {d['rd'] / 1e9:8.2f} {d['wr'] / 1e9:8.2f}")
