"""ctypes wrapper of the C oracle (oracle/c/mh_oracle.c): same algorithm as oracle/mh.py 'lean' mode, ~100x faster.
Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

from . import forward

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libmh_oracle.so")


def build():
    subprocess.run(["make", "-C", os.path.join(_HERE, "c")], check=True, capture_output=True)


def _lib():
    if not os.path.isfile(_LIB):
        build()
    lib = C.CDLL(_LIB)
    dp, ip, fp, up = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_uint8)
    model = [dp, ip, ip, dp, dp, C.c_double, dp, dp, dp, dp]
    lib.mh_oracle_taped.restype = C.c_long
    lib.mh_oracle_taped.argtypes = model + [C.c_int, C.c_int, fp, fp, up, fp, up, dp, fp]
    lib.mh_oracle_forced.restype = C.c_long
    lib.mh_oracle_forced.argtypes = model + [C.c_int, C.c_int, fp, fp, up, fp, up, up, up, dp, fp]
    lib.mh_oracle_free.restype = C.c_long
    lib.mh_oracle_free.argtypes = model + [C.c_int, C.c_int, C.c_int, C.c_uint64, fp]
    lib.mh_oracle_ll_roi.restype = C.c_double
    lib.mh_oracle_ll_roi.argtypes = model[:8] + [C.c_int, C.c_double, C.c_double]
    return lib


class CModel:
    """Packs an oracle.logp.Model for the C routines."""

    def __init__(self, m):
        self.lib = _lib()
        act, nrow = forward.active_columns(m.t)
        self.a = dict(M=np.ascontiguousarray(m.M), ncol=np.ascontiguousarray(nrow, np.int32),
                      acol=np.ascontiguousarray(act, np.int32), t=np.ascontiguousarray(m.t), cr=np.ascontiguousarray(m.c_r),
                      y=np.ascontiguousarray(m.y), sig=np.ascontiguousarray(m.sigma_noise),
                      mu=np.ascontiguousarray(np.stack(m.mu)), P=np.ascontiguousarray(np.stack(m.P)))
        self.k2p = float(m.k2p)

    def _args(self, n=10):
        a = self.a
        d = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))
        i = lambda x: x.ctypes.data_as(C.POINTER(C.c_int))
        return [d(a["M"]), i(a["ncol"]), i(a["acol"]), d(a["t"]), d(a["cr"]), self.k2p, d(a["y"]), d(a["sig"]), d(a["mu"]), d(a["P"])][:n]

    def ll_roi(self, roi, dvr, r1):
        return self.lib.mh_oracle_ll_roi(*self._args(8), int(roi), float(dvr), float(r1))

    def run_taped(self, tape, n_tune, n_draws):
        n = n_tune + n_draws
        draws = np.empty((n, 2, 48), np.float32)
        acc = np.empty((n, 2, 48), np.uint8)
        delta = np.empty((n, 2, 48), np.float64)
        scale = np.empty((2, 48), np.float32)
        f = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
        u = lambda x: x.ctypes.data_as(C.POINTER(C.c_uint8))
        self.lib.mh_oracle_taped(*self._args(), n, n_tune, f(tape.normals), f(tape.logu), u(tape.rank), f(draws), u(acc),
                                 delta.ctypes.data_as(C.POINTER(C.c_double)), f(scale))
        return dict(draws=draws, accept=acc.astype(bool), delta=delta, scale=scale)

    def run_forced(self, tape, n_tune, n_draws, forced_draws):
        """Teacher-forced replay (same semantics as oracle.mh.run_chain(..., forced_draws=...))."""
        n = n_tune + n_draws
        fd = np.ascontiguousarray(forced_draws, np.float32)
        acc = np.empty((n, 2, 48), np.uint8); facc = np.empty((n, 2, 48), np.uint8); und = np.empty((n, 2, 48), np.uint8)
        delta = np.empty((n, 2, 48), np.float64)
        scale = np.empty((2, 48), np.float32)
        f = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
        u = lambda x: x.ctypes.data_as(C.POINTER(C.c_uint8))
        self.lib.mh_oracle_forced(*self._args(), n, n_tune, f(tape.normals), f(tape.logu), u(tape.rank), f(fd), u(acc), u(facc),
                                  u(und), delta.ctypes.data_as(C.POINTER(C.c_double)), f(scale))
        return dict(accept=acc.astype(bool), forced_accept=facc.astype(bool), undecidable=und.astype(bool), delta=delta, scale=scale)

    def run_free(self, n_chains, n_tune, n_draws, seed=0, keep=True, threads=None):
        """n_chains free-running chains (internal generator), spread over `threads` OS threads
        (ctypes releases the GIL during the C call)."""
        from concurrent.futures import ThreadPoolExecutor
        n = n_tune + n_draws
        threads = min(threads or os.cpu_count(), n_chains)
        draws = np.empty((n_chains, n, 2, 48), np.float32) if keep else None
        bounds = np.linspace(0, n_chains, threads + 1).astype(int)

        def work(k):
            lo, hi = int(bounds[k]), int(bounds[k + 1])
            if hi <= lo:
                return 0
            ptr = draws[lo:].ctypes.data_as(C.POINTER(C.c_float)) if keep else None
            return self.lib.mh_oracle_free(*self._args(), hi - lo, n, n_tune, int(seed) + 7919 * lo, ptr)
        with ThreadPoolExecutor(threads) as ex:
            nacc = sum(ex.map(work, range(threads)))
        return draws, nacc
