#!/usr/bin/env python
"""Basic-block summary of one kernel launch from an ncu report: share of samples / executed warp instructions per block of SASS
with a common execution count, plus the instruction mix of the hottest blocks.
usage: python scripts/ncu_blocks.py prof.ncu-rep <launch index> [min share %]"""
import csv, subprocess, sys, collections, io
rep, skip = sys.argv[1], int(sys.argv[2])
minshare = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(skip), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
S, IE = ix['# Samples'], ix['Instructions Executed']
data = [r for r in rows[2:] if r and r[0].startswith('0x')]
half = len(data)//2
if half and [r[0] for r in data[:half]] == [r[0] for r in data[half:]]: data = data[:half]
tot = sum(int(r[S]) for r in data); toti = sum(int(r[IE]) for r in data)
print(rows[0][1][:100]); print('instructions', len(data), 'samples', tot, 'warp-instr executed', toti)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
print('  '.join('%s %.1f%%' % (s[6:], 100*v/max(tot, 1)) for s, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
blocks = []; cur = None
for k, r in enumerate(data):
    ie = int(r[IE])
    if cur is None or ie != cur[2]: cur = [k, k, ie, 0, 0]; blocks.append(cur)
    cur[1] = k; cur[3] += int(r[S]); cur[4] += ie
for b in blocks:
    if 100*b[3]/tot >= minshare or 100*b[4]/toti >= minshare:
        ops = collections.Counter()
        for r in data[b[0]:b[1]+1]:
            t = r[1].strip().split(); op = t[1] if t[0].startswith('@') else t[0]
            ops[op.split('.')[0]] += 1
        mix = ' '.join('%s:%d' % (o, n) for o, n in ops.most_common(7))
        print('[%4d-%4d] n=%3d exec=%11d samples %5.1f%% instr %5.1f%% | %s' % (b[0], b[1], b[1]-b[0]+1, b[2], 100*b[3]/tot, 100*b[4]/toti, mix))
