#!/usr/bin/env python
"""Aggregate an ncu source-page CSV (ncu -i X --page source --csv --print-source sass) by
opcode and by code region: executed warp-instructions and stall samples."""
import signal
signal.signal(signal.SIGPIPE, signal.SIG_DFL)
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot_inst = tot_samp = 0
byop, sampop = collections.Counter(), collections.Counter()
regions = []
for n, r in enumerate(data):
    src = r[ix['Source']].strip()
    parts = src.split()
    op = parts[1] if src.startswith('@') else parts[0]
    op = op.split('.')[0]
    ie = int(r[ix['Instructions Executed']]); sm = int(r[ix['# Samples']])
    tot_inst += ie; tot_samp += sm
    byop[op] += ie; sampop[op] += sm
    regions.append((n, src, ie, sm))
print('total warp-inst', tot_inst, 'samples', tot_samp)
for op, c in byop.most_common(28):
    print('%-10s %12d %5.1f%%   samples %5.1f%%' % (op, c, 100 * c / tot_inst, 100 * sampop[op] / tot_samp))
print()
step = int(sys.argv[2]) if len(sys.argv) > 2 else 250
for k in range(0, len(regions), step):
    ch = regions[k:k + step]
    print(k, 'inst %5.1f%%' % (100 * sum(c[2] for c in ch) / tot_inst), 'samples %5.1f%%' % (100 * sum(c[3] for c in ch) / tot_samp), ch[0][1][:60])
