#!/usr/bin/env python
"""Dev tool: per-loop breakdown of an ncu source-page CSV (ncu -i X --page source --csv --print-source sass) of the sweep
kernel: executed warp-instructions, stall samples and top stall reasons of every loop nest (backward branches) and of the
straight-line code between them.  Usage: python tools/ncu_regions.py src.csv [min_share_pct]"""
import signal
signal.signal(signal.SIGPIPE, signal.SIG_DFL)
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
addr = [int(r[ix['Address']], 16) if 'Address' in ix else n * 16 for n, r in enumerate(data)]
base = addr[0]
src = [r[ix['Source']].strip() for r in data]
ie = [int(r[ix['Instructions Executed']]) for r in data]
sm = [int(r[ix['# Samples']]) for r in data]
stall_cols = [h for h in hdr if h.startswith('stall_') and not h.endswith('(Not Issued)')]
tot_i, tot_s = sum(ie), sum(sm)
# loops
spans = []
for n, s in enumerate(src):
    m = re.search(r'\bBRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)', s)
    if m:
        tgt = int(m.group(1), 16)
        if tgt in addr:
            k = addr.index(tgt)
            if k < n:
                spans.append((k, n))
spans.sort(key=lambda s: (s[0], -s[1]))
def report(lo, hi, label):
    i = sum(ie[lo:hi + 1]); s = sum(sm[lo:hi + 1])
    if 100 * i / tot_i < minshare and 100 * s / tot_s < minshare:
        return
    st = collections.Counter()
    for n in range(lo, hi + 1):
        for c in stall_cols:
            v = data[n][ix[c]]
            if v:
                st[c[6:]] += int(v)
    ops = collections.Counter()
    for n in range(lo, hi + 1):
        t = re.sub(r'^@!?U?P\d+\s+', '', src[n]).split()[0].split('.')[0]
        ops[t] += ie[n]
    top = ', '.join('%s %.0f%%' % (k, 100 * v / max(1, sum(st.values()))) for k, v in st.most_common(4))
    topo = ', '.join('%s %.0f%%' % (k, 100 * v / max(1, i)) for k, v in ops.most_common(6))
    print('%-34s [%5d..%5d] %4d instr  inst %5.1f%%  samples %5.1f%%  | %s | %s' % (label, lo, hi, hi - lo + 1, 100 * i / tot_i, 100 * s / tot_s, top, topo))
print('total warp-inst', tot_i, 'samples', tot_s)
for lo, hi in spans:
    depth = sum(1 for a, b in spans if a <= lo and hi <= b) - 1
    report(lo, hi, '  ' * depth + 'loop')
