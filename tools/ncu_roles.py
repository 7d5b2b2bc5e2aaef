"""Per-role (code segment between EXITs) stall summary of an ncu report (dev tool)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]; data = rows[2:]
isrc = hdr.index('Source'); isamp = hdr.index('# Samples'); iex = hdr.index('Instructions Executed')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[isamp] or 0) for r in data)
seg = []; cur = {'n': 0, 'samp': 0, 'st': {}, 'start': 0, 'inst': 0}
for k, r in enumerate(data):
    cur['n'] += 1; cur['samp'] += int(r[isamp] or 0); cur['inst'] += int(r[iex] or 0)
    for i in stalls:
        cur['st'][hdr[i][6:]] = cur['st'].get(hdr[i][6:], 0) + int(r[i] or 0)
    if r[isrc].strip().endswith('EXIT') or 'EXIT' in r[isrc].split():
        seg.append(cur); cur = {'n': 0, 'samp': 0, 'st': {}, 'start': k + 1, 'inst': 0}
seg.append(cur)
print('total samples', tot)
for s in seg:
    if s['samp'] < 50: continue
    top = sorted(s['st'].items(), key=lambda x: -x[1])[:6]
    print('seg@%d n=%d samples=%d (%.1f%%) warp-instr=%d' % (s['start'], s['n'], s['samp'], 100 * s['samp'] / tot, s['inst']), top)
