"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): total / average time per kernel name."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hi]; data = rows[hi + 1:]
ki = h.index('Kernel Name'); vi = h.index('Metric Value'); ui = h.index('Metric Unit')
tot = collections.OrderedDict(); cnt = collections.Counter()
for r in data:
    if len(r) <= vi:
        continue
    n = r[ki][:70]; v = float(r[vi].replace(',', ''))
    v *= {'ns': 1, 'us': 1e3, 'ms': 1e6}.get(r[ui], 1)
    tot[n] = tot.get(n, 0) + v; cnt[n] += 1
T = sum(tot.values())
for n, v in sorted(tot.items(), key=lambda x: -x[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("%-70s %6d launches %11.1f us total %9.1f us avg %5.1f%%" % (n, cnt[n], v / 1e3, v / 1e3 / cnt[n], 100 * v / T))
print("total %.1f us over %d launches" % (T / 1e3, len(data)))
