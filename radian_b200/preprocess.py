"""Drop-in for the reference's ``radian/preprocess.py``, running on the GPU.

``mad_normalise`` and ``get_windows`` keep the reference's names, arguments, results and
ValueErrors (preprocess.py:4-49; call sites basecall.py:78,83); the batched forms take many reads
per call.  Raw signals are the int16 arrays a fast5 file holds (``radian_b200.fast5.reads``).
Nothing here computes: medians, z-scores and windows come from radian_b200/csrc/preprocess.cu.
"""
from __future__ import annotations

import numpy as np

from . import _native
from ._native import lib

_MSG = {4: "Signal must not be empty to normalise", 5: "MAD is zero, issue with signal."}


def _device(device):
    from .decode import _current_device

    return _current_device() if device is None else int(device)


def mad_normalise_batch(signals, outlier_z_score, device=None):
    """-> (list of arrays or ValueError instances, one per read).  A read the reference would
    refuse (empty signal, zero MAD) yields the ValueError it raises instead of an array."""
    sigs = []
    for s in signals:
        s = np.asarray(s)
        if s.dtype != np.int16:
            raise TypeError("raw signals are int16 (fast5 Raw/Signal); got " + str(s.dtype))
        sigs.append(np.ascontiguousarray(s).reshape(-1))
    n = len(sigs)
    off = np.zeros(n + 1, dtype=np.int64)
    off[1:] = np.cumsum([s.size for s in sigs]) if n else 0
    flat = np.concatenate(sigs) if n and off[-1] else np.zeros(1, dtype=np.int16)
    out = np.zeros(max(int(off[-1]), 1), dtype=np.float64)
    as_int = np.zeros(max(n, 1), dtype=np.int32)
    status = np.zeros(max(n, 1), dtype=np.int32)
    if n:
        rc = lib.radian_normalise_batch_host(_native.np_ptr(flat), _native.np_ptr(off), n, float(outlier_z_score),
                                             int(isinstance(outlier_z_score, (int, np.integer))),
                                             _native.np_ptr(out), _native.np_ptr(as_int), _native.np_ptr(status),
                                             _device(device))
        _native.check(rc)
    res = []
    for r in range(n):
        if status[r]:
            res.append(ValueError(_MSG.get(int(status[r]), f"status {int(status[r])}")))
            continue
        part = out[off[r]:off[r + 1]]
        res.append(part.view(np.int64).copy() if as_int[r] else part.copy())
    return res


def mad_normalise(signal, outlier_z_score):
    """Same call, result and errors as preprocess.mad_normalise (preprocess.py:23-29)."""
    r = mad_normalise_batch([signal], outlier_z_score)[0]
    if isinstance(r, ValueError):
        raise r
    return r


def get_windows_batch(signals, window_size, step_size, device=None):
    """-> list of ``(windows (n_w, window_size), pad_end)``, one per signal (float64 or int64)."""
    sigs = []
    for s in signals:
        s = np.asarray(s)
        if s.dtype.itemsize != 8 or s.dtype.kind not in "fi":
            raise TypeError("windows are cut from normalised signals (float64 / int64); got " + str(s.dtype))
        sigs.append(np.ascontiguousarray(s).reshape(-1))
    n = len(sigs)
    off = np.zeros(n + 1, dtype=np.int64)
    off[1:] = np.cumsum([s.size for s in sigs]) if n else 0
    nw = np.zeros(max(n, 1), dtype=np.int64)
    pad = np.zeros(max(n, 1), dtype=np.int32)
    _native.check(lib.radian_windows_plan(_native.np_ptr(off), n, int(window_size), int(step_size),
                                          _native.np_ptr(nw), _native.np_ptr(pad)))
    woff = np.zeros(n + 1, dtype=np.int64)
    woff[1:] = np.cumsum(nw[:n])
    flat = np.concatenate([s.view(np.float64) for s in sigs]) if n and off[-1] else np.zeros(1, dtype=np.float64)
    out = np.zeros((max(int(woff[-1]), 1), int(window_size)), dtype=np.float64)
    if n:
        _native.check(lib.radian_windows_batch_host(_native.np_ptr(flat), _native.np_ptr(off), n, int(window_size),
                                                    int(step_size), _native.np_ptr(woff), _native.np_ptr(out),
                                                    _device(device)))
    return [(out[woff[r]:woff[r + 1]].view(sigs[r].dtype).copy(), int(pad[r])) for r in range(n)]


def get_windows(signal, window_size, step_size):
    """Same call, result and errors as preprocess.get_windows (preprocess.py:4-21)."""
    return get_windows_batch([signal], window_size, step_size)[0]
