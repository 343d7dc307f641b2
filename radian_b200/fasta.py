"""FASTA output of a decoded batch (radian/basecall.py:129-141): ``>{read_id}\\n{sequence[::-1]}\\n``.

The records of a whole batch are formed in one call of the C ABI (radian_fasta_records_host) straight
from the decoder's symbol arrays, instead of one Python string per read.
"""
from __future__ import annotations

import numpy as np

from . import _native
from ._native import lib


def pack_ids(ids):
    """list of str -> (uint8 array of all ids back to back, int64 offsets[n+1])."""
    enc = [s.encode("ascii") for s in ids]
    off = np.zeros(len(enc) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(e) for e in enc])
    return np.frombuffer(b"".join(enc), dtype=np.uint8).copy(), off


def format_records(ids, seq, seq_offsets, lengths, bases="ACGT"):
    """FASTA text (uint8 array, supports the buffer protocol: ``f.write(text)``) of reads whose symbols
    0..3 are at ``seq[seq_offsets[r] : seq_offsets[r] + lengths[r]]`` in decode order; ``ids`` is a
    list of str or the result of ``pack_ids``.  Sequences are reversed as at basecall.py:129."""
    id_bytes, id_off = ids if isinstance(ids, tuple) else pack_ids(ids)
    n = len(id_off) - 1
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    so = np.ascontiguousarray(seq_offsets, dtype=np.int64)
    ln = np.ascontiguousarray(lengths, dtype=np.int64)
    if len(ln) != n or len(so) < n:
        raise ValueError("ids, seq_offsets and lengths disagree on the number of reads")
    if len(bases) != 4:
        raise ValueError("bases must have 4 letters")
    out_off = np.zeros(n + 1, dtype=np.int64)
    out_off[1:] = np.cumsum((id_off[1:] - id_off[:-1]) + ln + 3)
    out = np.empty(int(out_off[-1]), dtype=np.uint8)
    _native.check(lib.radian_fasta_records_host(_native.np_ptr(seq), _native.np_ptr(so), _native.np_ptr(ln), n,
                                                _native.np_ptr(id_bytes), _native.np_ptr(id_off), bases.encode("ascii"),
                                                _native.np_ptr(out), _native.np_ptr(out_off)))
    return out
