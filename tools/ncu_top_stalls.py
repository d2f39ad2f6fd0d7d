#!/usr/bin/env python
"""Top instructions by a given stall reason from an ncu source-page CSV."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
reason = sys.argv[2] if len(sys.argv) > 2 else "long_sb"
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix['# Samples']]) for r in data)
col = ['stall_' + reason, 'stall_' + reason + ' (Not Issued)']
v = [(sum(int(r[ix[c]] or 0) for c in col if c in ix), n) for n, r in enumerate(data)]
print('total', reason, '%.1f%%' % (100 * sum(x for x, _ in v) / tot))
for x, n in sorted(v, reverse=True)[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
    r = data[n]
    print(n, '%.2f%%' % (100 * x / tot), r[ix['Instructions Executed']].rjust(9), r[ix['Source']].strip()[:90])
