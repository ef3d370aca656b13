"""Draws from numpy's legacy global stream, taken by the library's restatement of the generator.

The reference's jacknife sweep and ``replace_md`` (/root/reference/locator/locator.py:722-727, :258-261)
consume ``np.random.binomial`` draws by the million; their order fixes the indices a seed produces, so
they cannot be parallelised, but numpy spends most of each draw on broadcasting overhead.
``legacy_binomial`` copies the MT19937 state out of ``np.random``, lets ``loc_np_legacy_binomial`` take
the same draws (same recurrence, same inversion sampler, same libm calls) and puts the advanced state
back: callers see exactly what the numpy calls would have produced, including the stream position.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._cabi import lib


def legacy_binomial(n, p, reps=1):
    """uint8 [len(p), reps]: what ``[np.random.binomial(n, p_i, reps) for p_i in p]`` returns, drawn from
    (and advancing) numpy's global stream.  Falls back to numpy itself where the library declines
    (a (n, p) pair in numpy's BTPE branch, n > 255, a non-MT19937 global generator)."""
    p = np.ascontiguousarray(p, dtype=np.float64).reshape(-1)
    reps = int(reps)
    out = np.empty((p.size, reps), dtype=np.uint8)
    if out.size == 0:
        return out
    st = np.random.get_state()
    if st[0] == "MT19937":
        key = np.ascontiguousarray(st[1], dtype=np.uint32).copy()
        pos = C.c_int32(int(st[2]))
        rc = lib.loc_np_legacy_binomial(key.ctypes.data, C.byref(pos), int(n), p.ctypes.data, p.size, reps,
                                        out.ctypes.data)
        if rc == 0:
            np.random.set_state(("MT19937", key, int(pos.value), st[3], st[4]))
            return out
        if rc == 2:
            raise ValueError("p < 0, p > 1 or p is NaN")  # numpy's message for the same input
    vals = np.random.binomial(int(n), np.repeat(p, reps)).reshape(p.size, reps)  # same draws, numpy's pace
    return vals.astype(np.uint8)


def legacy_permutation(n):
    """int64 [n]: what ``np.random.permutation(n)`` returns, drawn from (and advancing) numpy's global stream."""
    n = int(n)
    st = np.random.get_state()
    if st[0] != "MT19937" or n < 0:
        return np.random.permutation(n)
    out = np.empty(n, dtype=np.int64)
    key = np.ascontiguousarray(st[1], dtype=np.uint32).copy()
    pos = C.c_int32(int(st[2]))
    if lib.loc_np_legacy_permutation(key.ctypes.data, C.byref(pos), n, out.ctypes.data) != 0:
        return np.random.permutation(n)
    np.random.set_state(("MT19937", key, int(pos.value), st[3], st[4]))
    return out


def legacy_choice_without_replacement(n, size):
    """``np.random.choice(n, size, replace=False)`` for an integer population: the head of a permutation."""
    n, size = int(n), int(size)
    if size > n or size < 0:
        return np.random.choice(n, size, replace=False)  # numpy's own error for the same request
    return legacy_permutation(n)[:size]
