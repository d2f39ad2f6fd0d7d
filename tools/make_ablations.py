#!/usr/bin/env python
"""Dev tool: build ABLATION variants of libpetmh.so (each removes one part of the sweep kernel; results are WRONG by
construction, only the timing is meaningful) into build/variants/, to be timed back to back with
tools/variant_probe.sh.  This is how DESIGN.md section 3 ("where the time goes") was measured.

  x1  no visit-order rounds (every proposal rejected)      x2  no record step (moments / draw storage)
  x3  no likelihood (cheap function of the accumulators)   x4  no M.e product
  x7  erfc factor never evaluated                          x8  no ex2 in M.e (fma instead)
  x9  no rsqrt in the likelihood                           x10 no LDS in the full M.e columns

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
    ("        if (!tuning) {\n            if (!TAPED && active) {",
     "        if (!tuning) {\n            if (!TAPED && active && PETMH_ABLATE != 2) {"),
    # x3
    ("#pragma unroll 1\n        for (int it = 0; it < K; it++) {   // (fully unrolling this loop was measured slower: code size)",
     "#if PETMH_ABLATE == 3\n        {\n            u64 t0 = 0ull, t1 = 0ull, t2 = 0ull;\n#pragma unroll\n"
     "            for (int pq = 0; pq < NPAIR; pq++) { t0 = fadd2(t0, acc0[pq]); t1 = fadd2(t1, acc1[pq]); t2 = fadd2(t2, acc2[pq]); }\n"
     "            float x, y;\n"
     "            unpack2(t0, x, y); v0 += (x + y) * 1e-3f * coef0 + r10;\n"
     "            unpack2(t1, x, y); v1 += (x + y) * 1e-3f * coef1 + r11;\n"
     "            unpack2(t2, x, y); v2 += (x + y) * 1e-3f * coef2 + r12;\n        }\n"
     "#pragma unroll 1\n        for (int it = 0; it < 0; it++) {\n#else\n"
     "#pragma unroll 1\n        for (int it = 0; it < K; it++) {   // (fully unrolling this loop was measured slower: code size)\n#endif"),
    # x4
    ("        {\n            const int n0 = c_cend[blk][PETMH_TRIANGLE ? 0 : 8];",
     "#if PETMH_ABLATE == 4\n#pragma unroll\n        for (int pq = 0; pq < NPAIR; pq++) { acc0[pq] = pack2(na0 * -30.f, na0 * -30.f); "
     "acc1[pq] = pack2(na1 * -30.f, na1 * -30.f); acc2[pq] = pack2(na2 * -30.f, na2 * -30.f); }\n        if (false)\n#endif\n"
     "        {\n            const int n0 = c_cend[blk][PETMH_TRIANGLE ? 0 : 8];"),
    ("#if PETMH_TRIANGLE\n        PETMH_PHASE(1)", "#if PETMH_TRIANGLE && PETMH_ABLATE != 4\n        PETMH_PHASE(1)"),
    # x7
    ("    if (__any_sync(0xffffffffu, !(zmin >= Z_CUT))) {   // NaN -> evaluate",
     "    if (PETMH_ABLATE != 7 && __any_sync(0xffffffffu, !(zmin >= Z_CUT))) {"),
    # x8
    ("        E##0 = ex2_approx(na0 * tc_); E##1 = ex2_approx(na1 * tc_); E##2 = ex2_approx(na2 * tc_);                \\",
     "        if (PETMH_ABLATE == 8) { E##0 = fmaf(na0, tc_, 1.0f); E##1 = fmaf(na1, tc_, 1.0f); E##2 = fmaf(na2, tc_, 1.0f); } else { \\\n"
     "        E##0 = ex2_approx(na0 * tc_); E##1 = ex2_approx(na1 * tc_); E##2 = ex2_approx(na2 * tc_); }               \\"),
    # x9
    ("    const u64 rsp = pack2(rsqrt_approx(s0), rsqrt_approx(s1));",
     "    const u64 rsp = PETMH_ABLATE == 9 ? pack2(fmaf(s0, -0.05f, 1.0f), fmaf(s1, -0.05f, 1.0f)) : pack2(rsqrt_approx(s0), rsqrt_approx(s1));"),
    # x10
    ("#define PETMH_LOADCOL(B, ptr)                                                                                    \\\n"
     "    B##0 = (ptr)[0]; B##1 = (ptr)[1]; B##2 = (ptr)[2]; B##3 = (ptr)[3]; B##4 = *reinterpret_cast<const float2*>((ptr) + 4);",
     "#define PETMH_LOADCOL(B, ptr)                                                                                    \\\n"
     "    if (PETMH_ABLATE == 10) { const float w_ = __int_as_float((int)(size_t)(ptr)); B##0 = make_float4(w_, w_, w_, w_); "
     "B##1 = B##0; B##2 = B##0; B##3 = B##0; B##4 = make_float2(w_, w_); } else { \\\n"
     "    B##0 = (ptr)[0]; B##1 = (ptr)[1]; B##2 = (ptr)[2]; B##3 = (ptr)[3]; B##4 = *reinterpret_cast<const float2*>((ptr) + 4); }"),
    ("constexpr int NROI = 48;", "#ifndef PETMH_ABLATE\n#define PETMH_ABLATE 0\n#endif\nconstexpr int NROI = 48;"),
]


def main():
    which = [a.lstrip("x") for a in sys.argv[1:]] or ["1", "2", "3", "4", "7", "8", "9", "10"]
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
