"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name, count / total / share (dev tool)."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = None; agg = collections.OrderedDict(); n = collections.Counter()
for r in rows:
    if r[0] == 'ID':
        hdr = r; continue
    if hdr is None:
        continue
    name = r[hdr.index('Kernel Name')][:70]; v = float(r[hdr.index('Metric Value')].replace(',', ''))
    unit = r[hdr.index('Metric Unit')]
    v = v / 1e3 if unit in ('ns', 'nsecond') else (v * 1e3 if unit in ('ms', 'msecond') else v)
    agg[name] = agg.get(name, 0) + v; n[name] += 1
tot = sum(agg.values())
for k, v in sorted(agg.items(), key=lambda x: -x[1]):
    print("%-72s n=%5d  %12.1f us  %5.1f%%  (%.1f us each)" % (k, n[k], v, 100 * v / tot, v / n[k]))
print("total us %.1f" % tot)
