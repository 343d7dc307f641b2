"""Readers for the fixtures written by oracle/make_golden.py (see its docstring)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class DecodeCase:
    __slots__ = ("name", "mat", "bw", "L", "tseed", "s_thr", "r_thr", "seq", "scores",
                 "n_lookup", "n_combine")

    def __repr__(self):
        return (f"<{self.name} T={self.mat.shape[0]} {self.mat.dtype} bw={self.bw} L={self.L} "
                f"thr=({self.s_thr},{self.r_thr})>")


def decode_cases(fname):
    path = os.path.join(GOLDEN, fname)
    if not os.path.exists(path):
        return []
    z = np.load(path)
    out = []
    so = 0
    for i, m in enumerate(z["meta"]):
        is64, off, T, bw, L, tseed, nseq, nl, ncomb = (int(x) for x in m)
        c = DecodeCase()
        c.name = f"{fname}[{i}]"
        src = z["post64"] if is64 else z["post32"]
        c.mat = np.ascontiguousarray(src[off:off + T]).reshape(T, 5)
        c.bw, c.L, c.tseed = bw, L, tseed
        s, r = z["thr"][i]
        c.s_thr = None if np.isnan(s) else float(s)
        c.r_thr = None if np.isnan(r) else float(r)
        c.seq = z["seq"][so:so + nseq]
        so += nseq
        sc = z["scores"][i]
        c.scores = sc[~np.isnan(sc)]
        c.n_lookup, c.n_combine = nl, ncomb
        out.append(c)
    return out


def assembly_cases():
    z = np.load(os.path.join(GOLDEN, "assembly.npz"))
    out = []
    ci = ri = o32 = o64 = 0
    for S, n, rows, T, is64 in z["meta"]:
        lens = z["chunk_lens"][ci:ci + n]
        ci += n
        mats = []
        r = ri
        for ln in lens:
            mats.append(z["chunks"][r:r + ln].reshape(ln, 5))
            r += ln
        ri += rows
        if is64:
            ref = z["out64"][o64:o64 + T]
            o64 += T
        else:
            ref = z["out32"][o32:o32 + T]
            o32 += T
        out.append((int(S), mats, ref))
    return out


_TABLES = {}


def table(L, seed):
    from radian_b200 import synth

    if (L, seed) not in _TABLES:
        if len(_TABLES) > 4:
            _TABLES.clear()
        _TABLES[(L, seed)] = synth.make_table(L, seed)
    return _TABLES[(L, seed)]


def preprocess_cases():
    """-> list of dicts: signal (int16), outlier (python int or float), error (0 ok, 1 empty, 2 MAD
    zero), result (float64 or int64 array), windows: list of (W, S, n_windows, pad, sum, checksum)."""
    z = np.load(os.path.join(GOLDEN, "preprocess.npz"))
    off = z["offsets"]
    out = []
    for i in range(len(off) - 1):
        o = float(z["outlier"][i])
        bits = z["result_bits"][off[i]:off[i + 1]]
        out.append({
            "signal": z["signal"][off[i]:off[i + 1]],
            "outlier": int(o) if z["outlier_is_int"][i] else o,
            "error": int(z["error"][i]),
            "result": bits.copy() if z["result_is_int"][i] else bits.view(np.float64).copy(),
            "windows": [tuple(w[1:]) for w in z["windows"] if int(w[0]) == i],
        })
    return out


def window_checksum(w):
    return float(w.sum()), float((w * np.arange(1, w.size + 1).reshape(w.shape) % 7).sum())
