#!/usr/bin/env python
"""Dev tool: build ABLATION variants of libpetmh.so (each removes one part of the sweep kernel; results are WRONG by
construction, only the timing is meaningful) into build/variants/, to be timed back to back with
tools/variant_probe.sh.  This is how DESIGN.md section 3 ("where the time goes") was measured.

  x1  no visit-order rounds (every proposal rejected)      x2  no record step (moments / draw storage)
  x3  no likelihood (cheap function of the accumulators)   x4  no operator columns (A T(s): accumulators = R1 c_r only)
  x7  erfc factor never evaluated                          x9  no rsqrt in the likelihood
  x11 no Philox / Box-Muller (cheap hash instead)

The patches are textual and assert that their anchors still exist in petmh_device.cuh.
Usage: python tools/make_ablations.py [x1 x3 ...]   (default: all)
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "pet_posterior_distribution_b200", "csrc")
OUT = os.path.join(ROOT, "build", "variants")
TMP = os.path.join(ROOT, "build", "ablation_src")

PATCHES = [
    # x1
    ("            uint32_t last = 0u;                                       // keys <= last are decided\n",
     "            uint32_t last = 0u;                                       // keys <= last are decided\n"
     "#if PETMH_ABLATE == 1\n            key[0] = key[1] = key[2] = 0u;\n#endif\n"),
    # x2
    ("            if (!TAPED && active) {\n                // one base address per array;",
     "            if (!TAPED && active && PETMH_ABLATE != 2) {\n                // one base address per array;"),
    # x3
    ("        v0 += block_loglik(acc0, sCc + rowoff0 + blk * RSTRIDE, sYcc + rowoff0 + blk * RSTRIDE);\n"
     "        v1 += block_loglik(acc1, sCc + rowoff1 + blk * RSTRIDE, sYcc + rowoff1 + blk * RSTRIDE);\n"
     "        v2 += block_loglik(acc2, sCc + rowoff2 + blk * RSTRIDE, sYcc + rowoff2 + blk * RSTRIDE);\n",
     "#if PETMH_ABLATE == 3\n        {\n            u64 t0 = 0ull, t1 = 0ull, t2 = 0ull;\n#pragma unroll\n"
     "            for (int pq = 0; pq < NPAIR; pq++) { t0 = fadd2(t0, acc0[pq]); t1 = fadd2(t1, acc1[pq]); t2 = fadd2(t2, acc2[pq]); }\n"
     "            float x, y;\n"
     "            unpack2(t0, x, y); v0 += (x + y) * 1e-3f;\n"
     "            unpack2(t1, x, y); v1 += (x + y) * 1e-3f;\n"
     "            unpack2(t2, x, y); v2 += (x + y) * 1e-3f;\n        }\n#else\n"
     "        v0 += block_loglik(acc0, sCc + rowoff0 + blk * RSTRIDE, sYcc + rowoff0 + blk * RSTRIDE);\n"
     "        v1 += block_loglik(acc1, sCc + rowoff1 + blk * RSTRIDE, sYcc + rowoff1 + blk * RSTRIDE);\n"
     "        v2 += block_loglik(acc2, sCc + rowoff2 + blk * RSTRIDE, sYcc + rowoff2 + blk * RSTRIDE);\n#endif\n"),
    # x4: skip every operator column but the last (R1 c_r)
    ("            PETMH_LOADCOL(ma, Ap)\n            {   // columns 2, 3",
     "            PETMH_LOADCOL(ma, Ap)\n#if PETMH_ABLATE == 4\n"
     "            Ap += (n0 - 2) * (RSTRIDE / 4);\n            PETMH_LOADCOL(ma, Ap + 2 * (RSTRIDE / 4))\n            PETMH_FIRSTCOL(ma, a)\n"
     "            if (false)\n#endif\n            {   // columns 2, 3"),
    ("#pragma unroll 1\n            for (int c = 4; c < n0; c += 2) {",
     "#pragma unroll 1\n            for (int c = 4; c < (PETMH_ABLATE == 4 ? 0 : n0); c += 2) {"),
    ("            PETMH_LOADCOL(mb, Ap + (RSTRIDE / 4))\n            PETMH_FULLCOL(ma, cs)                        // column 1: T'_1 = coef s",
     "#if PETMH_ABLATE != 4\n            PETMH_LOADCOL(mb, Ap + (RSTRIDE / 4))\n            PETMH_FULLCOL(ma, cs)                        // column 1: T'_1 = coef s"),
    ("            PETMH_FULLCOL(ma, a)                         // R1 c_r (kinetic_model.py:157)",
     "            PETMH_FULLCOL(ma, a)                         // R1 c_r (kinetic_model.py:157)\n#endif"),
    # x7
    ("    if (__any_sync(0xffffffffu, !(zmin >= Z_CUT))) {   // NaN -> evaluate",
     "    if (PETMH_ABLATE != 7 && __any_sync(0xffffffffu, !(zmin >= Z_CUT))) {"),
    # x9
    ("    const float r0 = rsqrt_approx(s0), r1 = rsqrt_approx(s1);",
     "    const float r0 = PETMH_ABLATE == 9 ? fmaf(s0, -0.05f, 1.0f) : rsqrt_approx(s0), r1 = PETMH_ABLATE == 9 ? fmaf(s1, -0.05f, 1.0f) : rsqrt_approx(s1);"),
    # x11
    ("    uint32_t ci = (uint32_t)i, glo = (uint32_t)gid;\n    asm volatile(\"\" : \"+r\"(ci), \"+r\"(glo));",
     "    uint32_t ci = (uint32_t)i, glo = (uint32_t)gid;\n    asm volatile(\"\" : \"+r\"(ci), \"+r\"(glo));\n"
     "#if PETMH_ABLATE == 11\n    { const uint32_t hsh = (ci * 2654435761u) ^ ((uint32_t)(2 * sweep + b) * 40503u) ^ glo;\n"
     "      nrm = (float)(int)(hsh & 0xffffu) * 3.0e-5f - 1.0f; logu = -(float)(hsh >> 20) * 1.0e-3f; key = (hsh & 0xffffffc0u) | 0x80000000u | (uint32_t)i; return; }\n#endif"),
    ("constexpr int NROI = 48;", "#ifndef PETMH_ABLATE\n#define PETMH_ABLATE 0\n#endif\nconstexpr int NROI = 48;"),
]


def main():
    which = [a.lstrip("x") for a in sys.argv[1:]] or ["1", "2", "3", "4", "7", "9", "11"]
    shutil.rmtree(TMP, ignore_errors=True)
    os.makedirs(os.path.join(TMP, "pet_posterior_distribution_b200"), exist_ok=True)
    shutil.copytree(SRC, os.path.join(TMP, "pet_posterior_distribution_b200", "csrc"),
                    ignore=shutil.ignore_patterns("build.log"))
    shutil.copytree(os.path.join(ROOT, "include"), os.path.join(TMP, "include"))
    dev = os.path.join(TMP, "pet_posterior_distribution_b200", "csrc", "petmh_device.cuh")
    s = open(dev).read()
    for old, new in PATCHES:
        assert s.count(old) >= 1, "anchor not found (petmh_device.cuh changed): %r" % old[:70]
        s = s.replace(old, new, 1)
    open(dev, "w").write(s)
    os.makedirs(OUT, exist_ok=True)
    procs = []
    for n in which:
        cmd = ["nvcc", "-DPETMH_ABLATE=" + n, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
               "-Xcompiler", "-fPIC", "-shared", "-o", os.path.join(OUT, "libpetmh_x%s.so" % n), "petmh.cu"]
        procs.append((n, subprocess.Popen(cmd, cwd=os.path.dirname(dev), stderr=subprocess.PIPE, text=True)))
    for n, p in procs:
        err = p.communicate()[1]
        if p.returncode:
            sys.exit("ablation x%s failed to build:\n%s" % (n, err[-2000:]))
        print("built build/variants/libpetmh_x%s.so" % n)
    shutil.copy(os.path.join(ROOT, "pet_posterior_distribution_b200", "libpetmh.so"), os.path.join(OUT, "libpetmh_cur.so"))


if __name__ == "__main__":
    main()
