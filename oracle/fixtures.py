"""Shared helpers for seeded test inputs.  TEST INFRASTRUCTURE ONLY (used by tests/ and gen_golden.py).

Weights: the reference initialises with np.random.random(shape) / 100 (r_learning.py:136-149) and
writes files as float32 arrays grouped by weight_signature (r_learning.py:151-158).  The fixtures use
the same distribution from a legacy-seeded MT19937 stream (stable across numpy versions), rounded to
float32 so that the float64 reference, the float64/float32 oracle and the float32 device path all
start from bit-identical values.
"""
import numpy as np

SIGNATURE = {2: (24,), 3: (52,), 4: (17,), 5: (17, 4), 6: (17, 4, 12)}
GROUP_SIZE = {2: (256,), 3: (4096,), 4: (65536,), 5: (65536, 1048576), 6: (65536, 1048576, 7529536)}


def init_weights32(n, seed):
    """list of float32 arrays in the reference's weight-file layout."""
    rs = np.random.RandomState(seed)
    return [(rs.random_sample((d, s)) / 100).astype(np.float32) for d, s in zip(SIGNATURE[n], GROUP_SIZE[n])]


def flat(arrays):
    """weight-file arrays -> one flat vector in table order (table i at oracle.table_offsets(n)[i])."""
    return np.concatenate([np.asarray(a).reshape(-1) for a in arrays])


def unflat(n, vec):
    out, o = [], 0
    for d, s in zip(SIGNATURE[n], GROUP_SIZE[n]):
        out.append(np.asarray(vec[o:o + d * s]).reshape(d, s))
        o += d * s
    return out


def flat_from_ref_lists(weights):
    """reference in-memory weights (list of lists of float) -> flat float64 vector."""
    return np.concatenate([np.asarray(t, dtype=np.float64) for t in weights])


def sparse_diff(base, new):
    idx = np.nonzero(base != new)[0].astype(np.int64)
    return idx, new[idx]


def apply_sparse(base, idx, val):
    out = base.copy()
    out[idx] = val
    return out


def random_boards(m, seed, p_empty=0.3, max_exp=11):
    """BASELINE config-5 synthetic distribution: each cell empty with p=0.3 else uniform in 1..11."""
    rs = np.random.RandomState(seed)
    e = rs.randint(1, max_exp + 1, size=(m, 4, 4)).astype(np.int32)
    e[rs.random_sample((m, 4, 4)) < p_empty] = 0
    return e


def edge_boards():
    """hand-picked edge cases: empty, full-no-move, full-with-merges, 15s (overflow), ragged lines."""
    b = [
        np.zeros((4, 4)),
        [[1, 2, 1, 2], [2, 1, 2, 1], [1, 2, 1, 2], [2, 1, 2, 1]],            # game over
        [[1, 2, 1, 2], [2, 1, 2, 1], [1, 2, 1, 2], [2, 1, 2, 2]],            # one pair left
        [[1, 1, 1, 1], [1, 1, 1, 1], [1, 1, 1, 1], [1, 1, 1, 1]],
        [[15, 15, 15, 15], [14, 14, 14, 14], [0, 0, 0, 0], [1, 0, 0, 1]],    # overflow rows (left/right only)
        [[15, 0, 0, 0], [15, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0]],          # overflow in columns (up/down)
        [[1, 1, 2, 2], [0, 0, 0, 0], [3, 0, 3, 0], [1, 2, 3, 4]],            # SURVEY golden
        [[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12], [13, 14, 15, 0]],      # SURVEY feature golden
        [[0, 0, 0, 1], [0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0]],
        [[13, 14, 15, 13], [14, 13, 14, 15], [15, 14, 13, 14], [13, 15, 14, 13]],
        [[2, 2, 2, 0], [2, 0, 2, 2], [0, 2, 2, 2], [2, 2, 0, 2]],
    ]
    return np.array(b, dtype=np.int32)
