"""B200-native batched Metropolis-Hastings posterior sampler for SRTM2 PET kinetics.

Drop-in for ONE path of yanisdjebra/PET_posterior_distribution: its MCMC baseline
(`mcmc.py` over `kinetic_model.SRTM2`).  Host code is Python (as the reference's), the
compute path is hand-written sm_100a CUDA behind the C ABI in include/petmh.h.
No CPU fallback: importing without the built library raises ImportError.
"""
from ._lib import LIB_PATH, N_COORD, N_FRAMES, N_ROI, N_STATS, STAT_NAMES, PetmhError  # noqa: F401
from .sampler import MHSampler  # noqa: F401

__all__ = ["MHSampler", "PetmhError", "N_ROI", "N_FRAMES", "N_COORD", "N_STATS", "STAT_NAMES", "LIB_PATH"]
