"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/radian_oracle.c.

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports it.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libradian_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "radian_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libradian_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.radian_oracle_beam_search.restype = ctypes.c_int
        _lib.radian_oracle_beam_search_batch.restype = ctypes.c_int
        _lib.radian_oracle_assemble.restype = ctypes.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def beam_search(mat, beam_width, table=None, L=0, s_thr=0.0, r_thr=0.0, topk=8):
    """-> (symbols uint8[n], scores float64[<=topk], n_final, (n_lookup, n_combine))."""
    mat = np.ascontiguousarray(mat)
    assert mat.dtype in (np.float32, np.float64)
    T = mat.shape[0]
    if T:
        assert mat.shape[1] == 5
    if table is not None:
        table = np.ascontiguousarray(table, dtype=np.float64)
        assert table.shape == (4 ** L, 4)
    cap = max(T, 1)
    seq = np.zeros(cap, dtype=np.uint8)
    n = ctypes.c_int64(0)
    scores = np.full(topk, np.nan)
    nfin = ctypes.c_int(0)
    cnt = np.zeros(2, dtype=np.uint64)
    rc = lib().radian_oracle_beam_search(
        _p(mat), ctypes.c_int(mat.dtype == np.float64), ctypes.c_int64(T), ctypes.c_int(beam_width),
        _p(table), ctypes.c_int(L), ctypes.c_double(s_thr or 0.0), ctypes.c_double(r_thr or 0.0),
        _p(seq), ctypes.c_int64(cap), ctypes.byref(n), _p(scores), ctypes.c_int(topk),
        ctypes.byref(nfin), _p(cnt))
    if rc:
        raise RuntimeError(f"oracle beam_search rc={rc}")
    k = min(topk, nfin.value)
    return seq[:n.value].copy(), scores[:k].copy(), nfin.value, (int(cnt[0]), int(cnt[1]))


def beam_search_batch(post, frame_offsets, beam_width, table=None, L=0, s_thr=0.0, r_thr=0.0,
                      threads=1):
    """Decode a concatenated batch on `threads` host threads (ctypes releases the GIL).
    -> (list of uint8 arrays, scores float64[n], counters uint64[n,2])."""
    post = np.ascontiguousarray(post)
    fo = np.ascontiguousarray(frame_offsets, dtype=np.int64)
    n = len(fo) - 1
    if table is not None:
        table = np.ascontiguousarray(table, dtype=np.float64)
    so = np.zeros(n + 1, dtype=np.int64)
    so[1:] = np.cumsum(np.maximum(fo[1:] - fo[:-1], 1))
    seq = np.zeros(int(so[-1]), dtype=np.uint8)
    ln = np.zeros(n, dtype=np.int64)
    sc = np.zeros(n, dtype=np.float64)
    cnt = np.zeros((n, 2), dtype=np.uint64)
    L_ = lib()

    def run(r0, r1):
        rc = L_.radian_oracle_beam_search_batch(
            _p(post), ctypes.c_int(post.dtype == np.float64), _p(fo), ctypes.c_int64(r0),
            ctypes.c_int64(r1), ctypes.c_int(beam_width), _p(table), ctypes.c_int(L),
            ctypes.c_double(s_thr or 0.0), ctypes.c_double(r_thr or 0.0), _p(seq), _p(so), _p(ln),
            _p(sc), _p(cnt))
        if rc:
            raise RuntimeError(f"oracle batch rc={rc}")

    if threads <= 1 or n <= 1:
        run(0, n)
    else:
        # longest reads first, one read per task, so the pool stays balanced
        order = np.argsort(-(fo[1:] - fo[:-1]))
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda r: run(int(r), int(r) + 1), order))
    return [seq[so[i]:so[i] + ln[i]].copy() for i in range(n)], sc, cnt


def assemble(matrices, step):
    """assemble_matrices restatement -> np.ndarray (T,5) float32 or float64."""
    lens = np.array([len(m) for m in matrices], dtype=np.int32)
    nonempty = [np.ascontiguousarray(m, dtype=np.float32).reshape(-1, 5) for m in matrices if len(m)]
    chunks = np.concatenate(nonempty) if nonempty else np.zeros((0, 5), np.float32)
    T = 0
    f64 = False
    for k, ln in enumerate(lens):
        if ln:
            f64 |= k * step < T  # overlaps an earlier chunk's rows
            T = max(T, k * step + int(ln))
    out = np.zeros((T, 5), dtype=np.float64 if f64 else np.float32)
    t_out = ctypes.c_int64(0)
    rc = lib().radian_oracle_assemble(_p(chunks), _p(lens), ctypes.c_int(len(lens)), ctypes.c_int(step),
                                      _p(out), ctypes.c_int(f64), ctypes.c_int64(T), ctypes.byref(t_out))
    if rc == -4:
        raise IndexError("list index out of range")
    if rc:
        raise RuntimeError(f"oracle assemble rc={rc}")
    return out


def stitch(fragments):
    """simple_assembly + argmax + index2base restatement -> (consensus str, votes int32[4, length]).
    Raises IndexError where the reference does."""
    code = {"A": 0, "C": 1, "G": 2, "T": 3, "a": 0, "c": 1, "g": 2, "t": 3}
    n = len(fragments)
    off = np.zeros(n + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(f) for f in fragments])
    sym = np.array([code[ch] for f in fragments for ch in f], dtype=np.uint8)
    if sym.size == 0:
        sym = np.zeros(1, dtype=np.uint8)
    cap = 1000 * (n + 1)
    votes = np.zeros((4, cap), dtype=np.int32)
    cons = np.zeros(cap, dtype=np.uint8)
    out_len = ctypes.c_int64(0)
    L_ = lib()
    L_.radian_oracle_stitch.restype = ctypes.c_int
    rc = L_.radian_oracle_stitch(_p(sym), _p(off), ctypes.c_int(n), _p(votes), ctypes.c_int64(cap), _p(cons),
                                 ctypes.byref(out_len))
    if rc == -6:
        raise IndexError("index out of bounds for the vote buffer")
    if rc:
        raise RuntimeError(f"oracle stitch rc={rc}")
    ln = out_len.value
    return "".join("ACGT"[s] for s in cons[:ln]), votes[:, :ln].copy()


def mad_normalise(signal, outlier_z_score):
    """preprocess.mad_normalise restatement for int16 raw signals -> float64 or int64 array.
    Raises ValueError with the reference's messages."""
    sig = np.ascontiguousarray(signal, dtype=np.int16)
    n = sig.shape[0]
    out = np.zeros(max(n, 1), dtype=np.float64)
    is_int = ctypes.c_int(0)
    L_ = lib()
    L_.radian_oracle_mad_normalise.restype = ctypes.c_int
    rc = L_.radian_oracle_mad_normalise(_p(sig), ctypes.c_int64(n), ctypes.c_double(float(outlier_z_score)),
                                        ctypes.c_int(isinstance(outlier_z_score, (int, np.integer))), _p(out),
                                        ctypes.byref(is_int))
    if rc == -7:
        raise ValueError("Signal must not be empty to normalise")
    if rc == -8:
        raise ValueError("MAD is zero, issue with signal.")
    if rc:
        raise RuntimeError(f"oracle mad_normalise rc={rc}")
    return out[:n].view(np.int64).copy() if is_int.value else out[:n].copy()


def get_windows(signal, window_size, step_size):
    """preprocess.get_windows restatement -> (windows (n, W), pad_end)."""
    if step_size <= 0:
        raise ValueError("Step size must be > 0")
    if step_size > window_size:
        raise ValueError("Step size must be <= window size")
    sig = np.asarray(signal)
    pad = ctypes.c_int(0)
    L_ = lib()
    L_.radian_oracle_windows.restype = ctypes.c_int64
    nw = L_.radian_oracle_windows(ctypes.c_int64(sig.shape[0]), ctypes.c_int(window_size), ctypes.c_int(step_size),
                                  ctypes.byref(pad))
    out = np.zeros((nw, window_size), dtype=sig.dtype)
    for w in range(nw):
        part = sig[w * step_size:w * step_size + window_size]
        out[w, :len(part)] = part
    return out, pad.value
