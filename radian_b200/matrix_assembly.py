"""Drop-in for the reference's ``radian/matrix_assembly.py`` (assemble_matrices), on the GPU.

Same signature and result as matrix_assembly.py:6-10, including its quirks: the first chunk
covering a timestep wins (np.add's result is discarded at :52), rows covered more than once are
L1-normalised in float64 by sklearn's ``normalize`` (:53), and the result is float64 exactly
when at least one row was normalised (np.asarray promotion at :44).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _native
from ._native import lib


def _flatten(batch):
    """list (reads) of list (chunks) of (rows,5) arrays -> chunks, chunk_row_offsets, read_chunk_ranges"""
    rows = []
    cro = [0]
    rcr = [0]
    for mats in batch:
        for m in mats:
            m = np.asarray(m)
            if m.size == 0:
                m = np.zeros((0, 5), np.float32)
            if m.ndim != 2 or m.shape[1] != 5:
                raise ValueError(f"chunk matrices must be (rows, 5), got {m.shape}")
            rows.append(np.ascontiguousarray(m, dtype=np.float32))
            cro.append(cro[-1] + m.shape[0])
        rcr.append(len(cro) - 1)
    chunks = np.concatenate(rows) if rows and cro[-1] else np.zeros((0, 5), np.float32)
    return chunks, np.asarray(cro, np.int64), np.asarray(rcr, np.int64)


def assemble_batch(batch, step_size, device=None):
    """Assemble many reads in one launch.  ``batch``: list of per-read lists of chunk matrices.
    Returns a list of (T_r, 5) arrays; all float64 if any row of the batch is covered by more
    than one chunk... per read the dtype follows the reference (float32 for reads without overlap)."""
    from .decode import _current_device

    device = _current_device() if device is None else int(device)
    n = len(batch)
    if n == 0:
        return []
    chunks, cro, rcr = _flatten(batch)
    rows = np.zeros(n, dtype=np.int64)
    any_ov = ctypes.c_int(0)
    maxrows = ctypes.c_int32(0)
    _native.check(lib.radian_assemble_plan(_native.np_ptr(cro), _native.np_ptr(rcr), n, int(step_size),
                                           _native.np_ptr(rows), ctypes.byref(any_ov), ctypes.byref(maxrows)))
    oro = np.zeros(n + 1, dtype=np.int64)
    oro[1:] = np.cumsum(rows)
    f64 = bool(any_ov.value)
    out = np.zeros((int(oro[-1]), 5), dtype=np.float64 if f64 else np.float32)
    if oro[-1]:
        _native.check(lib.radian_assemble_batch_host(_native.np_ptr(chunks), _native.np_ptr(cro), _native.np_ptr(rcr),
                                                     _native.np_ptr(oro), n, int(step_size), _native.np_ptr(out),
                                                     int(f64), device))
    res = []
    for r in range(n):
        m = out[oro[r]:oro[r + 1]]
        if f64:
            # per-read dtype as the reference: float32 unless one of ITS rows was normalised
            one = ctypes.c_int(0)
            lib.radian_assemble_plan(_native.np_ptr(cro), _native.np_ptr(rcr[r:r + 2].copy()), 1, int(step_size),
                                     None, ctypes.byref(one), None)
            if not one.value:
                m = m.astype(np.float32)
        res.append(m)
    return res


def assemble_matrices(matrices, step_size):
    """Same as the reference's assemble_matrices (matrix_assembly.py:6-10)."""
    out = assemble_batch([list(matrices)], step_size)[0]
    if out.shape[0] == 0:
        return np.asarray([])  # collapse_vstack of an empty stack (matrix_assembly.py:37-44)
    return out


class AssemblePlan:
    """Host-side planning of a batch (row counts, float64 promotion, offsets) done once and kept on
    the device, so that the assembly itself is a single asynchronous kernel launch."""

    def __init__(self, chunk_row_offsets, read_chunk_ranges, step_size, device):
        import torch

        cro = np.ascontiguousarray(chunk_row_offsets, dtype=np.int64)
        rcr = np.ascontiguousarray(read_chunk_ranges, dtype=np.int64)
        n = len(rcr) - 1
        rows = np.zeros(n, dtype=np.int64)
        any_ov = ctypes.c_int(0)
        maxrows = ctypes.c_int32(0)
        _native.check(lib.radian_assemble_plan(_native.np_ptr(cro), _native.np_ptr(rcr), n, int(step_size),
                                               _native.np_ptr(rows), ctypes.byref(any_ov), ctypes.byref(maxrows)))
        oro = np.zeros(n + 1, dtype=np.int64)
        oro[1:] = np.cumsum(rows)
        self.n_reads = n
        self.step = int(step_size)
        self.max_chunk_rows = int(maxrows.value)
        self.total_rows = int(oro[-1])
        self.f64 = bool(any_ov.value)
        self.rows_in = int(cro[-1])
        self.d_cro = torch.from_numpy(cro).to(device)
        self.d_rcr = torch.from_numpy(rcr).to(device)
        self.out_row_offsets = torch.from_numpy(oro).to(device)


def assemble_batch_device(chunks, chunk_row_offsets=None, read_chunk_ranges=None, step_size=None, plan=None,
                          out=None):
    """Resident variant: ``chunks`` (rows,5) float32 CUDA tensor; either the two host offset arrays
    + step (a plan is built) or a ready ``AssemblePlan``.  Returns (out CUDA tensor (sum T,5),
    out_row_offsets CUDA int64 tensor).  Runs on the current torch stream without synchronising."""
    import torch

    if plan is None:
        plan = AssemblePlan(chunk_row_offsets, read_chunk_ranges, step_size, chunks.device)
    if chunks.shape[0] != plan.rows_in:
        raise ValueError("chunks tensor does not match the plan")
    if out is None:
        out = torch.empty((plan.total_rows, 5), dtype=torch.float64 if plan.f64 else torch.float32,
                          device=chunks.device)
    stream = torch.cuda.current_stream(chunks.device).cuda_stream
    _native.check(lib.radian_assemble_batch_dev(chunks.data_ptr(), plan.d_cro.data_ptr(), plan.d_rcr.data_ptr(),
                                                plan.out_row_offsets.data_ptr(), plan.n_reads, plan.step,
                                                plan.max_chunk_rows, plan.total_rows, out.data_ptr(), int(plan.f64),
                                                ctypes.c_void_p(stream)))
    return out, plan.out_row_offsets
