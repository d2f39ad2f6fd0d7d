"""ctypes binding of libpetmh.so (the C ABI in include/petmh.h).

There is no CPU fallback: importing this module without the built CUDA library raises.
Build it with ``python -c "import __graft_entry__ as g; g.build()"`` (or
``make -C pet_posterior_distribution_b200/csrc``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PETMH_LIB") or os.path.join(_HERE, "libpetmh.so")   # (PETMH_LIB: kernel-variant probes, tools/variant_probe.sh)

N_ROI, N_FRAMES, N_COORD, N_STATS = 48, 54, 96, 8
STAT_NAMES = ("mean", "sd", "mcse_mean", "ess_bulk", "ess_tail", "r_hat", "accept_rate", "scaling")
EXT_NAMES = ("hdi_3%", "hdi_97%", "mcse_sd", "ess_sd")


class PetmhError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("petmh error %d: %s" % (code, msg))
        self.code = code


class Cfg(C.Structure):
    _fields_ = [("device", C.c_int32), ("n_chains", C.c_int32), ("max_tacs", C.c_int32),
                ("max_draws", C.c_int32), ("seed", C.c_uint64), ("tac_gid0", C.c_uint64)]


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "libpetmh.so not built (%s): this package has no CPU fallback -- run "
            "`python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    H = C.c_void_p
    dp, fp, u8p, u32p = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_uint32)
    sig = {
        "petmh_create": (C.c_int, [C.POINTER(Cfg), C.POINTER(H)]),
        "petmh_destroy": (None, [H]),
        "petmh_last_error": (C.c_char_p, [H]),
        "petmh_version": (C.c_int, []),
        "petmh_set_frames": (C.c_int, [H, dp, dp]),
        "petmh_set_prior": (C.c_int, [H, dp, dp, dp, dp]),
        "petmh_set_data": (C.c_int, [H, C.c_int, dp, dp, dp, dp]),
        "petmh_set_global_ids": (C.c_int, [H, C.c_int, C.POINTER(C.c_uint64), C.c_uint64, C.c_uint64]),
        "petmh_set_data_f32": (C.c_int, [H, C.c_int, fp, fp, fp, fp]),
        "petmh_interp1d_linear": (C.c_int, [H, C.c_int, dp, C.c_int, dp, dp, C.c_int, dp]),
        "petmh_continuous_convolution": (C.c_int, [H, C.c_int, dp, dp, dp, C.c_int, C.c_int, dp]),
        "petmh_time_exponential": (C.c_int, [H, C.c_int, dp, C.c_int, dp, dp]),
        "petmh_synth": (C.c_int, [H, C.c_int, C.c_uint64, dp, dp, C.c_double, dp]),
        "petmh_synth_set_test_rule": (C.c_int, [H, C.c_double, dp, dp, dp]),
        "petmh_synth_get": (C.c_int, [H, fp, dp, fp, fp, C.POINTER(C.c_int)]),
        "petmh_forward": (C.c_int, [H, C.c_int, dp, dp, dp]),
        "petmh_forward_srtm": (C.c_int, [H, C.c_int, dp, dp, dp, dp]),
        "petmh_loglik": (C.c_int, [H, C.c_int, dp, dp, dp, dp]),
        "petmh_get_operator": (C.c_int, [H, C.c_int, dp]),
        "petmh_get_cheb_operator": (C.c_int, [H, C.c_int, fp, C.POINTER(C.c_int), dp, dp]),
        "petmh_philox_raw": (C.c_int, [H, C.c_uint64, C.c_uint32, C.c_uint32, u32p]),
        "petmh_reset": (C.c_int, [H]),
        "petmh_run": (C.c_int, [H, C.c_int, C.c_int, C.c_int]),
        "petmh_plan": (C.c_int, [H, C.c_int, C.c_int, C.c_int]),
        "petmh_advance": (C.c_int, [H, C.c_int]),
        "petmh_run_taped": (C.c_int, [H, C.c_int, C.c_int, C.c_int, C.c_int, fp, fp, u8p, fp, fp, u8p, fp]),
        "petmh_srtm_sample": (C.c_int, [H, dp, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, fp, fp, u8p, fp, fp, u8p, fp]),
        "petmh_n_stored": (C.c_int, [H]),
        "petmh_get_chains": (C.c_int, [H, fp, fp]),
        "petmh_get_summary": (C.c_int, [H, fp]),
        "petmh_get_summary_ext": (C.c_int, [H, fp]),
        "petmh_summary_device": (C.c_int, [H, C.c_void_p, C.c_void_p]),
        "petmh_summary_from_draws_device": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                                      C.c_void_p, C.c_void_p, C.c_void_p]),
        "petmh_summary_from_moments_device": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int),
                                                        C.POINTER(C.c_int), C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                                        C.c_void_p]),
        "petmh_export_summary_inputs": (C.c_int, [H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                                  C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int)]),
        "petmh_get_ess_cross_chain": (C.c_int, [H, fp]),
        "petmh_get_posterior_cov": (C.c_int, [H, dp, dp]),
        "petmh_get_state": (C.c_int, [H, fp, fp]),
        "petmh_set_state": (C.c_int, [H, fp, fp, C.c_int]),
        "petmh_checkpoint_bytes": (C.c_int64, [H]),
        "petmh_get_checkpoint": (C.c_int, [H, C.c_void_p, C.c_int64]),
        "petmh_set_checkpoint": (C.c_int, [H, C.c_void_p, C.c_int64]),
        "petmh_set_stream": (C.c_int, [H, C.c_void_p]),
        "petmh_synchronize": (C.c_int, [H]),
        "petmh_last_kernel_ms": (C.c_int, [H, fp, C.POINTER(C.c_int)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)      # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    return lib, tuple(sig)


lib, EXPORTS = _load()
