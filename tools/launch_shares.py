"""Per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python tools/launch_shares.py <launches.csv> [label]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        h, start = r, i
        break
ik, iv = h.index('Kernel Name'), h.index('Metric Value')
d = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) > iv:
        try:
            d.setdefault(r[ik], []).append(float(r[iv].replace(',', '')) / 1e3)
        except ValueError:
            pass
tot = sum(sum(v) for v in d.values())
n = sum(len(v) for v in d.values())
print("== %s" % (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print("   total %.1f us over %d launches (cold-cache, serialised: compare shares)" % (tot, n))
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print("   %-72s n=%4d %11.1f us %6.1f%%  avg %8.1f us" % (k[:72], len(v), sum(v), 100 * sum(v) / tot, sum(v) / len(v)))
