"""Drop-in for the chunk-mode stitching of the reference (``radian/sequence_assembly.py``),
running on the GPU.

``simple_assembly`` and ``index2base`` keep the reference's names, arguments and results
(sequence_assembly.py:19-48, 90-97; call site basecall.py:122-123).  A read's fragments only
depend on each other through a running position, so the batched entry point ``stitch_batch``
aligns all consecutive fragment pairs of all reads in one launch, places and votes in three more
(radian_b200/csrc/stitch.cu).  Nothing here computes; the library fails loudly without a GPU.
"""
from __future__ import annotations

import numpy as np

from . import _native
from ._native import lib

_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _pair in enumerate(("Aa", "Cc", "Gg", "Tt")):  # base_dict of sequence_assembly.py:42
    for _ch in _pair:
        _CODE[ord(_ch)] = _i


def _encode(fragment) -> np.ndarray:
    if isinstance(fragment, np.ndarray):
        return np.ascontiguousarray(fragment, dtype=np.uint8)
    raw = np.frombuffer(str(fragment).encode("latin-1", "replace"), dtype=np.uint8)
    sym = _CODE[raw]
    if sym.size and sym.max() > 3:
        raise KeyError(str(fragment)[int(np.argmax(sym > 3))])  # base_dict[base], sequence_assembly.py:47
    return sym


def stitch_flat(sym, frag_offsets, read_frag_ranges, device=None, return_votes=False):
    """Stitch on flat arrays: ``sym`` uint8 symbols 0..3 of all fragments back to back,
    ``frag_offsets`` (n_frags+1) into it, ``read_frag_ranges`` (n_reads+1) fragment indices per read.
    -> ``(consensus symbols uint8 back to back, offsets int64[n_reads+1])`` and, with
    ``return_votes``, the ``(columns, 4)`` int32 vote counts at the same offsets."""
    from .decode import _current_device

    device = _current_device() if device is None else int(device)
    sym = np.ascontiguousarray(sym, dtype=np.uint8)
    foff = np.ascontiguousarray(frag_offsets, dtype=np.int64)
    rfr = np.ascontiguousarray(read_frag_ranges, dtype=np.int64)
    n = len(rfr) - 1
    if n <= 0:
        return (np.zeros(0, np.uint8), np.zeros(1, np.int64)) + ((np.zeros((0, 4), np.int32),) if return_votes else ())
    # a slot as large as the sum of the read's fragment lengths always suffices
    slots = np.zeros(n + 1, dtype=np.int64)
    slots[1:] = np.cumsum(np.maximum(foff[rfr[1:]] - foff[rfr[:-1]], 1))
    seq = np.zeros(int(slots[-1]), dtype=np.uint8)
    ln = np.zeros(n, dtype=np.int64)
    status = np.zeros(n, dtype=np.int32)
    votes = np.zeros((int(slots[-1]), 4), dtype=np.int32) if return_votes else None
    if sym.size == 0:
        sym = np.zeros(1, dtype=np.uint8)
    rc = lib.radian_stitch_batch_host(_native.np_ptr(sym), _native.np_ptr(foff), _native.np_ptr(rfr), n,
                                      _native.np_ptr(seq), _native.np_ptr(slots), _native.np_ptr(ln),
                                      _native.np_ptr(status), _native.np_ptr(votes), device)
    _native.check(rc)
    off = np.zeros(n + 1, dtype=np.int64)
    off[1:] = np.cumsum(ln)
    idx = np.repeat(slots[:-1] - off[:-1], ln) + np.arange(int(off[-1]), dtype=np.int64)
    return (seq[idx], off) + ((votes[idx],) if return_votes else ())


def stitch_device(frag_sym, frag_start, frag_len, read_frag_ranges, long_pair_scratch_ints=0):
    """Resident variant on torch CUDA tensors, asynchronous on the current torch stream:
    ``frag_sym`` uint8, ``frag_start`` / ``frag_len`` int64 (n_frags) -- e.g. ``res.seq``,
    ``res.seq_offsets[:-1]``, ``res.lengths`` of a chunk-mode ``decode_batch_device`` call --
    ``read_frag_ranges`` int64 (n_reads+1).  -> (seq uint8, out_offsets int64[n_reads+1], lengths
    int64, status int32); every read's slot is the sum of its fragment lengths."""
    import ctypes

    import torch

    dev = frag_sym.device
    n_reads = read_frag_ranges.numel() - 1
    n_frags = frag_len.numel()
    csum = torch.zeros(n_frags + 1, dtype=torch.int64, device=dev)
    csum[1:] = torch.cumsum(frag_len, 0)
    slots = csum[read_frag_ranges]
    out_offsets = (slots - slots[0]).contiguous()
    total = int(out_offsets[-1].item())
    seq = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
    lengths = torch.empty(n_reads, dtype=torch.int64, device=dev)
    status = torch.empty(n_reads, dtype=torch.int32, device=dev)
    nbytes = lib.radian_stitch_workspace_bytes(n_frags, total, int(long_pair_scratch_ints))
    ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    _native.check(lib.radian_stitch_batch_dev(
        frag_sym.data_ptr(), frag_start.data_ptr(), frag_len.data_ptr(), read_frag_ranges.data_ptr(), n_reads,
        n_frags, seq.data_ptr(), out_offsets.data_ptr(), total, lengths.data_ptr(), status.data_ptr(), None,
        int(long_pair_scratch_ints), ws.data_ptr(), ws.numel(), ctypes.c_void_p(stream)))
    return seq, out_offsets, lengths, status


def stitch_batch(fragment_lists, device=None, return_votes=False, bases="ACGT"):
    """Consensus string of every read from its chunk-mode fragments (strings over ACGT, or uint8
    arrays of symbols 0..3 as the decoder returns them).  With ``return_votes`` also the
    ``(4, length)`` vote counts of every read, the array the reference's simple_assembly returns."""
    n = len(fragment_lists)
    frags = [[_encode(f) for f in fl] for fl in fragment_lists]
    rfr = np.zeros(n + 1, dtype=np.int64)
    rfr[1:] = np.cumsum([len(fl) for fl in frags]) if n else 0
    flat = [f for fl in frags for f in fl]
    foff = np.zeros(len(flat) + 1, dtype=np.int64)
    foff[1:] = np.cumsum([f.size for f in flat]) if flat else 0
    sym = np.concatenate(flat) if flat and foff[-1] else np.zeros(0, dtype=np.uint8)
    res = stitch_flat(sym, foff, rfr, device, return_votes)
    seq, off = res[0], res[1]
    lut = np.frombuffer(bases.encode("ascii"), dtype=np.uint8)
    out = [lut[seq[off[r]:off[r + 1]]].tobytes().decode("ascii") for r in range(n)]
    if return_votes:
        return out, [res[2][off[r]:off[r + 1]].T.astype(np.float64) for r in range(n)]
    return out


def simple_assembly(bpreads):
    """Same call and result as the reference (sequence_assembly.py:19-39): the ``(4, length)``
    float64 vote counts of one read's fragments."""
    return stitch_batch([list(bpreads)], return_votes=True)[1][0]


def index2base(read) -> str:
    """sequence_assembly.py:90-97: symbols 0..3 -> string over ACGT."""
    return "".join("ACGT"[int(x)] for x in read)
