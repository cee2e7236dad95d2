"""Interleaved row-band partition of a frame across ranks.

Row y belongs to rank ((y // band_rows) % n_ranks); every rank renders its rows into a compact buffer
(own rows in ascending order) and rank 0 scatters the gathered buffers back (c2rt_deinterleave on the
GPU; `scatter_rows` is the same index map in numpy, used by the CPU-only multi-rank tests)."""
import numpy as np


def owned_rows(height, rank, n_ranks, band_rows):
    y = np.arange(height)
    return y[(y // band_rows) % n_ranks == rank]


def rows_owned(height, rank, n_ranks, band_rows):
    return int(owned_rows(height, rank, n_ranks, band_rows).size)


def rows_padded(height, n_ranks, band_rows):
    """Rows of the largest per-rank compact buffer (rank 0 always owns the most)."""
    return rows_owned(height, 0, n_ranks, band_rows)


def scatter_rows(gathered, height, n_ranks, band_rows):
    """gathered: [n_ranks, rows_padded, ...] -> frame [height, ...]"""
    out = np.empty((height,) + gathered.shape[2:], gathered.dtype)
    for r in range(n_ranks):
        rows = owned_rows(height, r, n_ranks, band_rows)
        out[rows] = gathered[r, : rows.size]
    return out
