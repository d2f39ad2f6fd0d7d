"""The C-ABI shared library loads and exports every symbol include/petmh.h declares
(no compute calls here: this suite runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "petmh.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(petmh_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from pet_posterior_distribution_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libpetmh.so does not export %s" % n
    assert set(names) == set(_lib.EXPORTS), set(names) ^ set(_lib.EXPORTS)
    assert lib.petmh_version() >= 100


def test_no_cpu_fallback():
    """Without an sm_100 device, create fails loudly (PETMH_ENODEVICE), never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pet_posterior_distribution_b200 import MHSampler, PetmhError
    with pytest.raises(PetmhError) as e:
        MHSampler()
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pet_posterior_distribution_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_null_handle_is_an_error_not_a_crash():
    """Every entry point that takes a handle answers PETMH_EINVAL (or 0 for the two size queries) to a NULL handle -- an
    error code, like the header promises, not a segmentation fault.  Run in a child process so that a regression fails
    this test instead of killing the suite."""
    import subprocess
    import sys
    code = r"""
import ctypes as C, sys
sys.path.insert(0, %r)
from pet_posterior_distribution_b200 import _lib
lib = C.CDLL(_lib.LIB_PATH)
skip = ("petmh_create", "petmh_destroy", "petmh_version", "petmh_last_error")
for n in _lib.EXPORTS:
    if n in skip:
        continue
    f = getattr(lib, n); f.restype = C.c_int
    rc = f(*([C.c_void_p(0)] * 11))
    assert rc == (0 if n in ("petmh_n_stored", "petmh_checkpoint_bytes") else -1), (n, rc)
lib.petmh_destroy(C.c_void_p(0))
lib.petmh_last_error.restype = C.c_char_p
lib.petmh_last_error(C.c_void_p(0))
print("ok")
""" % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (r.returncode, r.stdout[-300:], r.stderr[-600:])


def test_header_is_plain_c_and_links(tmp_path):
    """include/petmh.h is the boundary a C / cgo / cffi caller binds: it compiles as strict C99 (no C++, no torch types) and a
    C program links against libpetmh.so and calls it (petmh_version; petmh_create fails with PETMH_ENODEVICE without a GPU)."""
    import shutil
    import subprocess
    from pet_posterior_distribution_b200 import _lib
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include "petmh.h"\n'
                   'int main(void) {\n'
                   '    petmh_cfg cfg = {0, 4, 1, 0, 0, 0};\n'
                   '    petmh_t* h = NULL;\n'
                   '    int rc = petmh_create(&cfg, &h);\n'
                   '    printf("%d %d %s\\n", petmh_version(), rc, rc ? petmh_last_error(NULL) : "created");\n'
                   '    if (!rc) petmh_destroy(h);\n'
                   '    return 0;\n}\n')
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-l:" + os.path.basename(_lib.LIB_PATH), "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120, check=True).stdout.split(None, 2)
    assert int(out[0]) >= 100
    import torch
    if not torch.cuda.is_available():
        assert int(out[1]) == -2 and "no CPU fallback" in out[2]
    else:
        assert int(out[1]) == 0
