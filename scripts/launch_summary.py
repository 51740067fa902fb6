"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel totals of the LAST step and its sequence.
usage: launch_summary.py launches.csv [n_last]   (n_last = launches per step, default: half of the list)"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
seq = [(re.sub(r"\(.*", "", r[ki]).replace("void ", ""), float(r[vi].replace(",", "")) / 1e3) for r in rows[hi + 2:] if len(r) > vi]
seq = [s for s in seq if "int_peak" not in s[0]]
n = int(sys.argv[2]) if len(sys.argv) > 2 else len(seq) // 2
seq = seq[-n:]
tot, cnt = collections.defaultdict(float), collections.Counter()
for k, v in seq:
    tot[k] += v
    cnt[k] += 1
total = sum(tot.values())
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{k:36s} {cnt[k]:4d} {tot[k]:10.1f} us {100 * tot[k] / total:5.1f} %")
print(f"{'total':36s} {len(seq):4d} {total:10.1f} us")
if "-v" in sys.argv:
    for i, (k, v) in enumerate(seq):
        print(i, k, round(v, 1))
