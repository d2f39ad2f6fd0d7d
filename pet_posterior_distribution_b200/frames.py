"""The reference's 54-frame acquisition grid (sample_sim_data.py:29-85): six regular
segments of 10/15/30/60/120/300-second frames from 0 to 7200 s."""
import numpy as np

_SEGMENTS = ((60, 10), (180, 15), (360, 30), (840, 60), (1800, 120), (7200, 300))
MK_HALF_T = 109.8   # sample_sim_data.py:96 (18F half-life, minutes)


def acquisition_time_frames():
    """(54, 2) array of [start, end] in minutes, as sample_sim_data.py:29-82 defines it."""
    edges = [0.0]
    for end, step in _SEGMENTS:
        while edges[-1] < end:
            edges.append(edges[-1] + step)
    e = np.asarray(edges, np.float64)
    return 1 / 60 * np.stack([e[:-1], e[1:]], axis=1)


def frame_grid():
    """(time_vector, dt): frame end times and durations in minutes (sample_sim_data.py:84-85)."""
    f = acquisition_time_frames()
    return f[:, 1].copy(), f[:, 1] - f[:, 0]
