"""CPU ORACLE -- test infrastructure, not product code.

``linear_sum_assignment`` restated (see ``lsap.c`` for the provenance note): the shortest
augmenting path algorithm of Crouse (2016) as scipy.optimize.linear_sum_assignment
implements it, called by the reference at lib/modeling/matcher.py:93 and :158.

Two implementations with identical results: a pure-Python one (small cases, always
available) and the C one in ``oracle/_build/liblsap_oracle.so`` (``make -C oracle``), which
``linear_sum_assignment`` prefers when it has been built.  Both are pinned against
scipy 1.18.1 outputs in ``tests/golden/lsap_cases.npz``.
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblsap_oracle.so")
_lib = None


def _load():
    global _lib
    if _lib is None and os.path.exists(_SO):
        lib = ctypes.CDLL(_SO)
        lib.lsap_solve.restype = ctypes.c_int
        lib.lsap_solve.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_void_p]
        _lib = lib
    return _lib


def build(force: bool = False) -> str:
    """Compile lsap.c with gcc (used by ``__graft_entry__.build()`` and the tests)."""
    import subprocess
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "lsap.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lsap_python(cost) -> Tuple[np.ndarray, np.ndarray]:
    """Pure-Python restatement; mirrors lsap.c line for line."""
    cost = np.asarray(cost, dtype=np.float64)
    nr, nc = cost.shape
    if nr == 0 or nc == 0:
        return np.empty(0, np.int64), np.empty(0, np.int64)
    transpose = nc < nr
    if transpose:
        cost = cost.T.copy()
        nr, nc = nc, nr
    if np.isnan(cost).any() or np.isneginf(cost).any():
        raise ValueError("matrix contains invalid numeric entries")
    u = [0.0] * nr
    v = [0.0] * nc
    path = [-1] * nc
    col4row = [-1] * nr
    row4col = [-1] * nc
    c = cost.tolist()
    for cur in range(nr):
        best = 0.0
        remaining = [nc - t - 1 for t in range(nc)]
        n_rem = nc
        SR = [False] * nr
        SC = [False] * nc
        spc = [math.inf] * nc
        i, sink = cur, -1
        while sink == -1:
            pick, lowest = -1, math.inf
            SR[i] = True
            ci, ui = c[i], u[i]
            for t in range(n_rem):
                j = remaining[t]
                r = best + ci[j] - ui - v[j]
                if r < spc[j]:
                    path[j] = i
                    spc[j] = r
                if spc[j] < lowest or (spc[j] == lowest and row4col[j] == -1):
                    lowest = spc[j]
                    pick = t
            best = lowest
            if best == math.inf:
                raise ValueError("cost matrix is infeasible")
            j = remaining[pick]
            if row4col[j] == -1:
                sink = j
            else:
                i = row4col[j]
            SC[j] = True
            n_rem -= 1
            remaining[pick] = remaining[n_rem]
        u[cur] += best
        for i2 in range(nr):
            if SR[i2] and i2 != cur:
                u[i2] += best - spc[col4row[i2]]
        for j2 in range(nc):
            if SC[j2]:
                v[j2] -= best - spc[j2]
        j = sink
        while True:
            i2 = path[j]
            row4col[j] = i2
            col4row[i2], j = j, col4row[i2]
            if i2 == cur:
                break
    if transpose:
        pairs = [(j, row4col[j]) for j in range(nc) if row4col[j] != -1]
        a = np.array([p[0] for p in pairs], np.int64)
        b = np.array([p[1] for p in pairs], np.int64)
    else:
        a = np.arange(nr, dtype=np.int64)
        b = np.array(col4row, np.int64)
    return a, b


def lsap_c(cost) -> Tuple[np.ndarray, np.ndarray]:
    lib = _load()
    if lib is None:
        raise RuntimeError("oracle/_build/liblsap_oracle.so not built (make -C oracle)")
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    nr, nc = cost.shape
    k = min(nr, nc)
    a = np.empty(k, np.int64)
    b = np.empty(k, np.int64)
    rc = lib.lsap_solve(nr, nc, cost.ctypes.data, a.ctypes.data, b.ctypes.data)
    if rc == -2:
        raise ValueError("matrix contains invalid numeric entries")
    if rc == -1:
        raise ValueError("cost matrix is infeasible")
    return a, b


def linear_sum_assignment(cost) -> Tuple[np.ndarray, np.ndarray]:
    """Accepts any real 2-D array (fp32 costs are promoted to fp64 exactly, as scipy does)."""
    return lsap_c(cost) if _load() is not None else lsap_python(cost)
