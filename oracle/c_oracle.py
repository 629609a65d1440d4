"""ORACLE -- TEST INFRASTRUCTURE ONLY.  ctypes access to oracle/_build/libvsum_oracle.so
(`make -C oracle`), the plain-C restatement in `oracle/vsum_oracle.c`."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libvsum_oracle.so")
_lib = None


def build() -> str:
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.vsum_oracle_pairwise_sum_f32.restype = C.c_float
        L.vsum_oracle_pairwise_sum_f32.argtypes = [C.c_void_p, C.c_int64]
        L.vsum_oracle_capacity.restype = C.c_int32
        L.vsum_oracle_capacity.argtypes = [C.c_int32]
        L.vsum_oracle_knapsack.restype = None
        L.vsum_oracle_knapsack.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.vsum_oracle_fscore.restype = C.c_double
        L.vsum_oracle_fscore.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                                         C.c_int, C.c_void_p]
        L.vsum_oracle_video.restype = C.c_double
        L.vsum_oracle_video.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                                        C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                                        C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def pairwise_sum_f32(a: np.ndarray) -> np.float32:
    a = np.ascontiguousarray(a, dtype=np.float32)
    return np.float32(lib().vsum_oracle_pairwise_sum_f32(_p(a), len(a)))


def knapsack(cap: int, wt, val) -> list[int]:
    wt = np.ascontiguousarray(wt, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64)
    sel = np.zeros(len(wt), dtype=np.uint8)
    lib().vsum_oracle_knapsack(int(cap), _p(wt), _p(val), len(wt), _p(sel))
    return [int(i) for i in np.nonzero(sel)[0]]


def fscore(summary, user_summary, method: str = "avg"):
    summary = np.ascontiguousarray(summary, dtype=np.int8)
    us = np.ascontiguousarray(user_summary, dtype=np.float32)
    per_user = np.zeros(us.shape[0], dtype=np.float64)
    f = lib().vsum_oracle_fscore(_p(summary), len(summary), _p(us), us.shape[0], us.shape[1],
                                 1 if method == "max" else 0, _p(per_user))
    return f, per_user


def video(scores, picks, n_frames, change_points, user_summary, method: str = "avg"):
    """Returns dict(f, val, wt, cap, selected, summary) for one video."""
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    picks = np.ascontiguousarray(picks, dtype=np.int32)
    cps = np.ascontiguousarray(change_points, dtype=np.int32)
    us = np.ascontiguousarray(user_summary, dtype=np.float32)
    S = cps.shape[0]
    val = np.zeros(S, dtype=np.float64)
    wt = np.zeros(S, dtype=np.int32)
    cap = C.c_int32(0)
    sel = np.zeros(S, dtype=np.uint8)
    summ = np.zeros(int(cps[-1, 1]) + 1, dtype=np.int8)
    f = lib().vsum_oracle_video(_p(scores), len(scores), _p(picks), len(picks), int(n_frames),
                                _p(cps), S, _p(us), us.shape[0], us.shape[1],
                                1 if method == "max" else 0, _p(val), _p(wt), C.byref(cap),
                                _p(sel), _p(summ))
    return dict(f=f, val=val, wt=wt, cap=int(cap.value), selected=sel, summary=summ)
