#!/usr/bin/env python
"""Dev tool: static SASS statistics of the sweep kernel's hot loop (size, instruction mix) for a built library.
Usage: python tools/sass_stats.py [path/to/libpetmh.so]"""
import collections, re, signal, subprocess, sys
signal.signal(signal.SIGPIPE, signal.SIG_DFL)
lib = sys.argv[1] if len(sys.argv) > 1 else "pet_posterior_distribution_b200/libpetmh.so"
K = "_ZN5petmh15mh_sweep_kernelILi0ELb0ELi0EEEvNS_11SweepParamsE"
out = subprocess.run(["cuobjdump", "-sass", "-fun", K, lib], capture_output=True, text=True, check=True).stdout
ins = []
for line in out.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
redux = [a for a, t in ins if "REDUX" in t]
spans = []
for a, t in ins:
    m = re.search(r"\bBRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        spans.append((int(m.group(1), 16), a))
sweep = max((s for s in spans if any(s[0] <= r <= s[1] for r in redux)), key=lambda s: s[1] - s[0])
hot = [t for a, t in ins if sweep[0] <= a <= sweep[1]]
print("kernel %d instr; hot loop %d instr = %.1f KB" % (len(ins), len(hot), len(hot) / 64))
cnt = collections.Counter()
for t in hot:
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    cnt[t.split()[0].split(".")[0] + ("." + t.split()[0].split(".")[1] if t.startswith("MUFU") else "")] += 1
print(", ".join("%s %d" % kv for kv in cnt.most_common(40)))
print("inner loops inside the hot loop (start, len):", [(hex(lo), (hi - lo) // 16 + 1) for lo, hi in spans if sweep[0] <= lo and hi <= sweep[1] and (lo, hi) != sweep])
