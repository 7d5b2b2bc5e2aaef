"""Rank SASS instructions of an ncu report by stall samples (dev tool). usage: ncu_top.py report.ncu-rep [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]; data = rows[2:]
ia = hdr.index('Address'); isrc = hdr.index('Source'); isamp = hdr.index('# Samples'); iex = hdr.index('Instructions Executed')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[isamp] or 0) for r in data)
print('total samples', tot, 'instructions', len(data))
for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:N]:
    s = int(r[isamp])
    st = sorted(((hdr[i], int(r[i] or 0)) for i in stalls if int(r[i] or 0) > 0), key=lambda x: -x[1])[:3]
    print(r[ia][-5:], '%5.1f%%' % (100 * s / tot), r[iex].rjust(10), r[isrc][:52].ljust(52), st)
agg = {}
for r in data:
    for i in stalls:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print(sorted(agg.items(), key=lambda x: -x[1])[:8])
# cumulative samples in address order, 40 buckets
cum = 0; bucket = max(1, len(data) // 40)
for k in range(0, len(data), bucket):
    s = sum(int(r[isamp] or 0) for r in data[k:k + bucket])
    ex = max(int(r[iex] or 0) for r in data[k:k + bucket])
    print('  [%4d..] %5.1f%%  maxexec %d  %s' % (k, 100 * s / tot, ex, data[k][isrc][:40]))
