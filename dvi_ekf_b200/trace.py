"""FilterTraj / save_trajectory output of a batched run: the per-step trace that
``BatchFilter.run(trace=...)`` fills (eskf_streams_t.trace_x) turned into the reference's 30-column rows
(dvi_ekf/models/trajectory/FilterTraj.py:12-69) and its text format (dvi_ekf/tools/files.py:68-82), so that the
engine's output can be diffed against the reference's kf_best_*.txt artefacts."""
from __future__ import annotations

import numpy as np

from .filter import FilterTraj, State, save_trajectory


def filter_traj_rows(t0: float, x0, t_imu, x_trace) -> np.ndarray:
    """Rows of one filter: the initial state at ``t0`` (Filter.__init__ appends it, Filter.py:92) followed by one
    row per IMU step; ``x_trace`` is that filter's [T,26] slice of the trace (update instants already hold the
    updated state, as FilterTraj.append_updated_states overwrites the last row)."""
    rows = [FilterTraj.row(t0, State.from_vector(np.asarray(x0, dtype=float)))]
    for t, x in zip(t_imu, x_trace):
        rows.append(FilterTraj.row(t, State.from_vector(x)))
    return np.array(rows)


def save_filter_traj(filename: str, t0: float, x0, t_imu, x_trace) -> np.ndarray:
    """Writes the kf_best_*.txt file of one filter of the batch; returns the rows."""
    rows = filter_traj_rows(t0, x0, t_imu, x_trace)
    save_trajectory(rows, filename)
    return rows
