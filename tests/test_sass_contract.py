"""Build-time performance contract of the sweep kernel, checked on the compiled library (no GPU needed):
the properties DESIGN.md section 3 relies on and that a careless edit silently loses.

* 128 registers -> two 256-thread CTAs per SM (16 warps/SM); a small stack frame (eval3 inlined: as a real call the
  ABI spills every value live across it, 232-byte frames that thrash the ~30 KB of L1 left beside the shared memory);
* the hot loop (one sweep = two block updates) fits the 32 KB L1.5 instruction cache: every measured variant that
  grew it beyond that lost 2-15 %;
* the instructions the design is built on are really there: packed FFMA2 for A.T(s) and the likelihood, MUFU for
  exp/rsqrt/rcp/lg2, REDUX for the visit-order rounds, 128-bit shared-memory loads of the packed operator.
"""
import os
import re
import shutil
import subprocess

import pytest

from pet_posterior_distribution_b200 import LIB_PATH

KERNEL = "_ZN5petmh15mh_sweep_kernelILi0ELb0ELi0EEEvNS_11SweepParamsE"   # mh_sweep_kernel<0, false, 0>
cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
pytestmark = pytest.mark.skipif(not os.path.exists(cuobjdump), reason="cuobjdump not available")


def _run(*args):
    return subprocess.run([cuobjdump, *args, LIB_PATH], capture_output=True, text=True, check=True).stdout


def test_registers_and_stack_of_the_sweep_kernel():
    out = _run("--dump-resource-usage")
    m = re.search(re.escape(KERNEL) + r":\s*\n\s*REG:(\d+) STACK:(\d+)", out)
    assert m, "sweep kernel not found in the library"
    regs, stack = int(m.group(1)), int(m.group(2))
    assert regs <= 128, "more than 128 registers: only one 256-thread CTA per SM would fit"
    assert stack <= 128, "stack frame %d B: eval3 no longer inlined, or new spills (see DESIGN.md section 3)" % stack


def _kernel_sass():
    out = _run("-sass", "-fun", KERNEL)
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    assert len(ins) > 1000
    return ins


def test_hot_loop_fits_the_instruction_cache_and_uses_the_intended_instructions():
    ins = _kernel_sass()
    # loops = backward branches; the sweep loop is the longest one that contains the REDUX of the rounds
    redux = [a for a, t in ins if "REDUX" in t]
    assert redux, "no REDUX: the visit-order rounds lost their warp reduction"
    spans = []
    for a, t in ins:
        m = re.search(r"\bBRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            spans.append((int(m.group(1), 16), a))
    sweep = max((s for s in spans if any(s[0] <= r <= s[1] for r in redux)), key=lambda s: s[1] - s[0])
    n_hot = (sweep[1] - sweep[0]) // 16 + 1
    assert n_hot * 16 <= 32 * 1024, "hot loop %d instructions = %.1f KB > 32 KB L1.5 instruction cache" % (n_hot, n_hot / 64)
    hot = [t for a, t in ins if sweep[0] <= a <= sweep[1]]

    def count(pat):
        return sum(1 for t in hot if re.search(pat, t))
    assert count(r"\bFFMA2\b") >= 120          # packed fp32x2 FMAs: Chebyshev-operator columns + likelihood
    assert count(r"\bMUFU\.EX2\b") >= 9 and count(r"\bMUFU\.RSQ\b") >= 18    # erfc factor 2^R(z); rsqrt of the model TAC
    assert count(r"\bMUFU\.LG2\b") >= 3
    assert count(r"\bLDS\.128\b") >= 20        # broadcast loads of the packed operator / per-ROI rows
    assert count(r"\bDFMA\b") >= 6             # fp64 prior bookkeeping in the rounds
    # spills stay outside the M.e column loop and the likelihood loop: no local-memory access in any inner loop that
    # contains FFMA2 but not the REDUX
    for lo, hi in spans:
        if (lo, hi) == sweep or any(lo <= r <= hi for r in redux):
            continue
        body = [t for a, t in ins if lo <= a <= hi]
        if sweep[0] <= lo and hi <= sweep[1] and any("FFMA2" in t for t in body) and len(body) < 600:
            assert not any(re.search(r"\b(LDL|STL)\b", t) for t in body), "local-memory spill inside an inner loop"
