"""Acquisition frame grid of the reference (sample_sim_data.py:29-85), restated from
its structure (Appendix A.1 of SURVEY.md), not copied: six regular segments."""
import numpy as np

# (segment end [s], frame length [s]) -- 6+8+6+8+8+18 = 54 frames, 0 .. 7200 s
_SEGMENTS = ((60, 10), (180, 15), (360, 30), (840, 60), (1800, 120), (7200, 300))


def frame_edges_seconds():
    edges = [0.0]
    for end, step in _SEGMENTS:
        while edges[-1] < end:
            edges.append(edges[-1] + step)
    return np.asarray(edges, dtype=np.float64)


def frame_grid():
    """Returns (time_vector, dt): frame END times and durations in minutes,
    exactly as sample_sim_data.py:84-85 forms them (1/60 * seconds)."""
    e = frame_edges_seconds()
    frames = 1 / 60 * np.stack([e[:-1], e[1:]], axis=1)
    return frames[:, 1].copy(), frames[:, 1] - frames[:, 0]


N_FRAMES = 54
N_ROI = 48
MK_HALF_T = 109.8  # sample_sim_data.py:96
