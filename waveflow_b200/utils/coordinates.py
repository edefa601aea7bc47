"""Coordinate helpers -- reference: utils/coordinates.py (inversion counts for the antisymmetry sign of sorted walkers)."""
from __future__ import annotations

import numpy as np


def get_num_inversion_count(coordinates) -> np.ndarray:
    """[batch, n] -> [batch] number of inversions (pairs i < j with x_i > x_j), i.e. the parity of the sorting permutation
    used as (-1)**count in utils/helpers.py:57-59 (utils/coordinates.py:38-48 counts them with a heap, one row at a time)."""
    x = np.asarray(coordinates)
    if x.ndim == 1:
        x = x[None]
    n = x.shape[1]
    iu, ju = np.triu_indices(n, k=1)
    return (x[:, iu] > x[:, ju]).sum(-1)
