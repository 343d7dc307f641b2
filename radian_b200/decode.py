"""Drop-in for the reference's ``radian/decode.py`` beam search, running on the GPU.

``beam_search`` keeps the reference's positional signature (decode.py:100-109) and return
value (the best labeling as a string in decode order, decode.py:207-212).  Because one read
per call cannot feed a B200, the batched entry points ``beam_search_batch`` (host arrays)
and ``decode_batch_device`` (resident torch CUDA tensors, asynchronous) are added; they
call the same kernel.  All arithmetic happens in libradian_b200.so; nothing here computes.
"""
from __future__ import annotations

import ctypes
import threading
import weakref
from dataclasses import dataclass

import numpy as np

from . import _native
from ._native import lib

N_BASES = 4  # decode.py:14


class RnaTable:
    """The RNA k-mer model resident in HBM: dense ``(4**L, 4)`` float64 rows indexed by the
    2-bit packed context (oldest symbol most significant, basecall.py:54-57), plus the row
    entropies the reference memoises in ``entr_cache`` (decode.py:86-90)."""

    def __init__(self, dense: np.ndarray, device: int = 0, present: "np.ndarray | None" = None):
        """``present``: optional (4**L,) mask of the contexts the model holds; a read whose search
        reaches an absent one fails with the reference's KeyError (decode.py:83)."""
        dense = np.ascontiguousarray(dense, dtype=np.float64)
        if dense.ndim != 2 or dense.shape[1] != 4:
            raise ValueError("table must have shape (4**L, 4)")
        L = 0
        while 4 ** L < dense.shape[0]:
            L += 1
        if 4 ** L != dense.shape[0] or L < 1:
            raise ValueError(f"table has {dense.shape[0]} rows, not a power of 4")
        h = ctypes.c_void_p()
        if present is not None:
            present = np.ascontiguousarray(present, dtype=np.uint8)
            if present.shape != (dense.shape[0],):
                raise ValueError("present must have one entry per context")
            if present.all():
                present = None
        _native.check(lib.radian_table_create_sparse(_native.np_ptr(dense), _native.np_ptr(present), L, int(device),
                                                     ctypes.byref(h)))
        self._h = h
        self.complete = present is None
        self.L = L
        self.device = int(device)

    @classmethod
    def from_dict(cls, lm: dict, len_context: int, device: int = 0) -> "RnaTable":
        """Dense copy of the dict built at basecall.py:50-57 ({tuple of ints: [pA,pC,pG,pT]}).
        A context missing from the dict is a KeyError in the reference when the search first
        visits it (decode.py:83), and so it is here: absent contexts are marked in the table and a
        read that reaches one raises KeyError(context)."""
        L = int(len_context)
        n = 4 ** L
        dense = np.full((n, 4), 0.25, dtype=np.float64)
        seen = np.zeros(n, dtype=bool)
        for ctx, dist in lm.items():
            if len(ctx) != L:
                continue
            idx = 0
            for c in ctx:
                idx = idx * 4 + int(c)
            dense[idx] = dist
            seen[idx] = True
        return cls(dense, device, present=seen)

    @classmethod
    def from_json(cls, path: str, device: int = 0, cache: bool = True) -> "RnaTable":
        """Load the reference's RNA model JSON ({"ACGT...": [pA,pC,pG,pT]}, basecall.py:50-57); the
        dense form is cached on disk (see ``load_rna_json``)."""
        dense, present = load_rna_json(path, cache=cache)
        return cls(dense, device, present=present)

    def entropies(self) -> np.ndarray:
        out = np.empty(4 ** self.L, dtype=np.float64)
        _native.check(lib.radian_table_entropies(self._h, _native.np_ptr(out)))
        return out

    def __bool__(self):
        return True

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib.radian_table_destroy(h)
            except Exception:
                pass


# Tables built from plain dicts, most recently used last.  Bounded: a dense table is 32 B x 4^L of
# HBM (128 MiB at L = 11), and plain dicts cannot be weak-referenced, so nothing else would ever
# evict them.  A hit must match the dict's identity, size and a sample of its rows; an in-place edit
# of other rows between two calls is not noticed -- pass an RnaTable to be explicit about lifetime.
def _cache_path(path: str) -> str:
    """Where the dense form of the model JSON at `path` is kept: next to the JSON if that directory
    is writable, else under $XDG_CACHE_HOME / ~/.cache; the name carries size and mtime of the JSON,
    so an edited model never meets a stale cache."""
    import hashlib
    import os

    st = os.stat(path)
    tag = f"{st.st_size:x}-{st.st_mtime_ns:x}"
    d = os.path.dirname(os.path.abspath(path))
    if os.access(d, os.W_OK):
        return os.path.join(d, f".{os.path.basename(path)}.{tag}.radian_dense.npz")
    root = os.path.join(os.environ.get("XDG_CACHE_HOME", os.path.join(os.path.expanduser("~"), ".cache")), "radian_b200")
    os.makedirs(root, exist_ok=True)
    return os.path.join(root, hashlib.sha1(os.path.abspath(path).encode()).hexdigest()[:16] + f"-{tag}.npz")


def load_rna_json(path: str, cache: bool = True):
    """The reference's RNA model JSON (basecall.py:47-57: {"ACGT...": [pA,pC,pG,pT]}, key = context,
    oldest base first) as ``(dense (4**L, 4) float64, present (4**L,) uint8)``.  Parsing the 4^11 keys of
    the real model takes many seconds of pure Python; the dense form is therefore written beside the JSON
    (uncompressed .npz) and a later start reads that instead (a fraction of a second)."""
    import json
    import os

    cp = _cache_path(path) if cache else None
    if cp and os.path.exists(cp):
        try:
            z = np.load(cp)
            return z["dense"], z["present"]
        except Exception:
            pass  # unreadable cache: rebuild it
    with open(path, "r") as f:
        raw = json.load(f)
    if not raw:
        raise ValueError(f"{path}: empty RNA model")
    L = len(next(iter(raw)))
    n = 4 ** L
    tr = str.maketrans("ACGTU", "01233")
    keys = [k for k in raw if len(k) == L]
    idx = np.fromiter((int(k.translate(tr), 4) for k in keys), dtype=np.int64, count=len(keys))
    dense = np.full((n, 4), 0.25, dtype=np.float64)
    present = np.zeros(n, dtype=np.uint8)
    dense[idx] = np.array([raw[k] for k in keys], dtype=np.float64).reshape(len(keys), 4)
    present[idx] = 1
    if cp:
        try:
            tmp = cp + f".{os.getpid()}.tmp.npz"
            np.savez(tmp, dense=dense, present=present)
            os.replace(tmp, cp)
        except OSError:
            pass  # read-only location: the next start parses again
    return dense, present


_TABLE_CACHE_MAX = 4
_table_cache: "dict[tuple, tuple]" = {}
_table_lock = threading.Lock()


def _resolve_table(lm, len_context, device):
    """lm truthiness follows decode.py:157,180: None / {} switch the model off."""
    if isinstance(lm, RnaTable):
        return lm
    if not lm:
        return None
    if isinstance(lm, str):
        # basecall.py:48-49 leaves --rna-model None as the *string* "None"; the reference then
        # dies with this TypeError at decode.py:83 as soon as a beam is long enough.
        raise TypeError("string indices must be integers, not 'tuple'")
    key = (id(lm), int(len_context), int(device))
    with _table_lock:
        hit = _table_cache.get(key)
        # id() of a dead dict can be reused by a new one: a hit also has to look like the same table
        mark = _fingerprint(lm, int(len_context))
        if hit is not None and hit[1] == mark:
            _table_cache[key] = _table_cache.pop(key)  # most recently used last
            return hit[0]
        t = RnaTable.from_dict(lm, len_context, device)
        _table_cache.pop(key, None)
        _table_cache[key] = (t, mark)
        while len(_table_cache) > _TABLE_CACHE_MAX:
            _table_cache.pop(next(iter(_table_cache)))  # the table is destroyed when its last user lets go
        try:
            weakref.finalize(lm, _table_cache.pop, key, None)
        except TypeError:
            pass  # plain dicts cannot be weak-referenced: the entry lives until it is pushed out
        return t


def _fingerprint(lm, L):
    """Size plus the rows of a few fixed contexts: cheap, and enough to tell two model dicts apart."""
    n = 4 ** L
    picks = sorted({(i * 0x9E3779B1) % n for i in range(64)} | {0, n - 1})
    rows = []
    for i in picks:
        ctx = tuple((i >> (2 * (L - 1 - j))) & 3 for j in range(L))
        row = lm.get(ctx) if hasattr(lm, "get") else None
        rows.append(None if row is None else tuple(float(x) for x in row))
    return (len(lm), tuple(rows))


def _current_device() -> int:
    try:
        import torch

        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except Exception:
        pass
    return 0


def trim_memory(device=None):
    """Return the device buffers the host entry points keep between calls to the driver."""
    _native.check(lib.radian_trim_memory(_current_device() if device is None else int(device)))


def beam_search_batch(mats, beam_width, lm=None, s_threshold=None, r_threshold=None, len_context=None,
                      bases="ACGT", device=None, return_details=False, return_symbols=False):
    """Decode a list of (T_i, 5) posterior matrices (all float32 or all float64) in one launch.
    Returns the list of decoded strings, or (strings, scores[n,2], counters[n,4]) with
    ``return_details`` (counters: lm reads, combine_dists calls, near-tie frames, 0; see
    include/radian_b200.h), or with ``return_symbols`` the compact ``(symbols uint8, offsets int64[n+1])``
    pair that ``sequence_assembly.stitch_flat`` takes."""
    if len(bases) != N_BASES:
        raise ValueError("this build decodes 4 bases + blank (chars == 5, decode.py:124-125)")
    if int(beam_width) < 1 or int(beam_width) > _native.MAX_BEAM_WIDTH:
        raise ValueError(f"beam_width must be in 1..{_native.MAX_BEAM_WIDTH}")
    device = _current_device() if device is None else int(device)
    table = _resolve_table(lm, len_context, device)
    n = len(mats)
    if n == 0:
        if return_symbols:
            return np.zeros(0, np.uint8), np.zeros(1, np.int64)
        return ([], np.zeros((0, 2)), np.zeros((0, 4), np.uint64)) if return_details else []
    # The reference decodes every matrix in its own dtype (float32 window / single-chunk matrices,
    # float64 assembled ones: matrix_assembly.py:36-53), and the float32 path differs in the S, p/S
    # and entropy arithmetic (decode.py:54-55, 67-76 under numpy promotion): a mixed batch is
    # therefore split by dtype, one launch each, and scattered back in the caller's order.
    arrs = []
    for m in mats:
        m = np.asarray(m)
        if m.ndim != 2 or m.shape[1] != N_BASES + 1:
            if m.size == 0:
                m = m.reshape(0, N_BASES + 1)
            else:
                raise ValueError(f"posterior matrix must be (T, 5), got {m.shape}")
        if m.dtype != np.float64:
            m = m.astype(np.float32, copy=False)
        arrs.append(np.ascontiguousarray(m))
    is64 = np.array([a.dtype == np.float64 for a in arrs])
    seqs = [None] * n
    score = np.zeros((n, 2), dtype=np.float64)
    cnt = np.zeros((n, 4), dtype=np.uint64) if return_details else None
    for want64 in (False, True):
        idx = np.flatnonzero(is64 == want64)
        if len(idx) == 0:
            continue
        sub, sc, c = _decode_host([arrs[i] for i in idx], np.float64 if want64 else np.float32, beam_width, table,
                                  s_threshold, r_threshold, len_context, device, return_details)
        for k, i in enumerate(idx):
            seqs[i] = sub[k]
        score[idx] = sc
        if cnt is not None:
            cnt[idx] = c
    if return_symbols:
        # compact symbols 0..3 of all reads back to back + offsets: what stitch_flat takes
        off = np.zeros(n + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(q) for q in seqs])
        return (np.concatenate(seqs) if n else np.zeros(0, np.uint8)), off
    lut = np.frombuffer(bases.encode("ascii"), dtype=np.uint8)
    out = [lut[q].tobytes().decode("ascii") for q in seqs]
    return (out, score, cnt) if return_details else out


def _decode_host(arrs, dt, beam_width, table, s_threshold, r_threshold, len_context, device, want_counters,
                 full_slots=False):
    """One radian_decode_batch_host_reads call on same-dtype matrices -> (symbol arrays, scores, counters)."""
    n = len(arrs)
    nf = np.array([a.shape[0] for a in arrs], dtype=np.int64)
    # one pointer per read: the library gathers the matrices itself, nothing is concatenated here
    ptrs = (ctypes.c_void_p * n)(*[a.ctypes.data if a.shape[0] else None for a in arrs])
    # output slots: a decoded read is far shorter than its frame count (a symbol needs a frame, real
    # reads take dozens); T/4 + 64 symbols, and the whole batch again with T if some read needs more
    so = np.zeros(n + 1, dtype=np.int64)
    so[1:] = np.cumsum(np.maximum(nf, 1) if full_slots else nf // 4 + 64)
    seq = np.empty(int(so[-1]), dtype=np.uint8)
    ln = np.zeros(n, dtype=np.int64)
    score = np.zeros((n, 2), dtype=np.float64)
    status = np.zeros(n, dtype=np.int32)
    cnt = np.zeros((n, 4), dtype=np.uint64) if want_counters else None
    rc = lib.radian_decode_batch_host_reads(
        ptrs, _native.np_ptr(nf), int(dt == np.float64), n, int(beam_width),
        table._h if table else None, int(len_context) if table else 0,
        float(s_threshold) if table else 0.0, float(r_threshold) if table else 0.0,
        _native.np_ptr(seq), _native.np_ptr(so), _native.np_ptr(ln), _native.np_ptr(score),
        _native.np_ptr(status), _native.np_ptr(cnt), device)
    if rc == _native.E_READ and not full_slots and (status == _native.READ_SEQ_OVERFLOW).any():
        return _decode_host(arrs, dt, beam_width, table, s_threshold, r_threshold, len_context, device, want_counters,
                            full_slots=True)
    if rc == _native.E_READ and (status == _native.READ_KEY_ERROR).any():
        bad = int(np.flatnonzero(status == _native.READ_KEY_ERROR)[0])
        L, ci = int(len_context), int(ln[bad])
        raise KeyError(tuple((ci >> (2 * (L - 1 - i))) & 3 for i in range(L)))  # decode.py:83
    if rc == _native.E_READ and (status == _native.READ_RANGE).any():
        bad = int(np.flatnonzero(status == _native.READ_RANGE)[0])
        raise FloatingPointError(
            f"read {bad}: beam scores spread over more than the float64 exponent range "
            "(the reference's log-domain scores have no such limit, decode.py:172-175); see INTEGRATION.md")
    _native.check(rc)
    return [seq[so[i]:so[i] + ln[i]] for i in range(n)], score, cnt


def beam_search(mat, bases, beam_width, lm, s_threshold, r_threshold, len_context, entr_cache):
    """Beam search decoder -- same signature and result as the reference (decode.py:100-212).

    ``entr_cache`` is accepted for compatibility and ignored: it is a pure memo of row
    entropies in the reference (decode.py:86-90); here they are computed once per table.
    """
    del entr_cache
    return beam_search_batch([mat], beam_width, lm, s_threshold, r_threshold, len_context, bases=bases)[0]


# --------------------------------------------------------------------------- resident path
@dataclass
class DeviceDecodeResult:
    seq: "object"          # uint8 CUDA tensor, symbols 0..3
    seq_offsets: "object"  # int64 CUDA tensor (n+1)
    lengths: "object"      # int64 CUDA tensor (n)
    scores: "object"       # float64 CUDA tensor (n, 2)
    status: "object"       # int32 CUDA tensor (n)
    counters: "object"     # uint64-as-int64 CUDA tensor (n, 4) or None
    table: "object" = None  # keeps the RnaTable alive while the launch that uses it is in flight
    workspace: "object" = None  # ... and the launch's workspace

    def strings(self, bases="ACGT"):
        seq = self.seq.cpu().numpy()
        so = self.seq_offsets.cpu().numpy()
        ln = self.lengths.cpu().numpy()
        lut = np.frombuffer(bases.encode("ascii"), dtype=np.uint8)
        return [lut[seq[so[i]:so[i] + ln[i]]].tobytes().decode("ascii") for i in range(len(ln))]


def _workspace(device: int, nbytes: int):
    """Queue head + back-pointer arenas of one launch.  Allocated per call from torch's caching
    allocator on the current stream (stream-ordered reuse: a later call on the same stream may get
    the same block back, a call on another stream never gets it while this launch can still be
    running), and kept alive by the result object."""
    import torch

    return torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{device}")


def decode_batch_device(post, frame_offsets, beam_width, table=None, s_threshold=0.0, r_threshold=0.0,
                        *, max_frames, order=None, seq_offsets=None, counters=False, out=None, arena_nodes=0):
    """Asynchronous decode of reads already resident in HBM (torch CUDA tensors), on the
    current torch stream.  ``post``: (sum T, 5) float32/float64; ``frame_offsets``: int64
    (n+1); ``order``: optional int32 processing order (longest first).  ``seq_offsets`` gives
    every read its output slot (default: T_r bytes, always enough).  Reads whose status is
    RADIAN_READ_TRIE_OVERFLOW (2) must be re-run with ``arena_nodes = 32 * (T + 1)``."""
    import torch

    dev = post.device.index
    n = frame_offsets.numel() - 1
    if seq_offsets is None:
        seq_offsets = torch.zeros(n + 1, dtype=torch.int64, device=post.device)
        seq_offsets[1:] = torch.cumsum(torch.clamp(frame_offsets[1:] - frame_offsets[:-1], min=1), 0)
        total = int(post.shape[0]) + n
    else:
        total = int(seq_offsets[-1].item())
    if out is None:
        out = DeviceDecodeResult(
            seq=torch.empty(total, dtype=torch.uint8, device=post.device),
            seq_offsets=seq_offsets,
            lengths=torch.empty(n, dtype=torch.int64, device=post.device),
            scores=torch.empty((n, 2), dtype=torch.float64, device=post.device),
            status=torch.empty(n, dtype=torch.int32, device=post.device),
            counters=torch.zeros((n, 4), dtype=torch.int64, device=post.device) if counters else None)
    nbytes = lib.radian_decode_workspace_bytes(dev, int(beam_width), int(n), int(max_frames), int(arena_nodes))
    if nbytes == 0:
        raise _native.RadianError(f"no CUDA device / bad beam width: {_native.last_error()}")
    # (a result object that is filled again brings its workspace along: same caller, same stream)
    ws = out.workspace if out.workspace is not None and out.workspace.numel() >= nbytes else _workspace(dev, nbytes)
    stream = torch.cuda.current_stream(post.device).cuda_stream
    rc = lib.radian_decode_batch_dev(
        post.data_ptr(), int(post.dtype == torch.float64), frame_offsets.data_ptr(), n,
        order.data_ptr() if order is not None else None, int(max_frames), int(post.shape[0]), int(beam_width),
        table._h if table is not None else None, table.L if table is not None else 0,
        float(s_threshold), float(r_threshold), out.seq.data_ptr(), out.seq_offsets.data_ptr(),
        out.lengths.data_ptr(), out.scores.data_ptr(), out.status.data_ptr(),
        out.counters.data_ptr() if out.counters is not None else None, int(arena_nodes),
        ws.data_ptr(), ws.numel(), ctypes.c_void_p(stream))
    _native.check(rc)
    out.table = table
    out.workspace = ws
    return out
