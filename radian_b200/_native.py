"""ctypes binding of libradian_b200.so (the C ABI declared in include/radian_b200.h).

There is no fallback: if the shared library is missing and cannot be built here (nvcc), the
import fails, and every compute call fails loudly when no CUDA device is present.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_void_p

from . import build as _build

OK, E_ARG, E_CUDA, E_CONTEXT, E_READ, E_GAP = 0, -1, -2, -3, -4, -5
READ_SEQ_OVERFLOW = 1  # RADIAN_READ_SEQ_OVERFLOW
READ_RANGE = 6  # RADIAN_READ_RANGE
READ_KEY_ERROR = 7  # RADIAN_READ_KEY_ERROR
MAX_BEAM_WIDTH = 128
MAX_CONTEXT = 13

EXPORTS = [
    "radian_last_error", "radian_version", "radian_device_count", "radian_trim_memory",
    "radian_table_create", "radian_table_create_sparse", "radian_table_destroy", "radian_table_context_len", "radian_table_entropies",
    "radian_decode_workspace_bytes", "radian_decode_batch_dev", "radian_decode_batch_host",
    "radian_decode_batch_host_reads",
    "radian_assemble_plan", "radian_assemble_batch_dev", "radian_assemble_batch_host",
    "radian_stitch_batch_host", "radian_stitch_batch_dev", "radian_stitch_workspace_bytes",
    "radian_normalise_batch_host", "radian_normalise_batch_dev", "radian_windows_plan",
    "radian_windows_batch_host", "radian_windows_batch_dev", "radian_fasta_records_host",
]


def _load():
    so = _build.SO
    if _build.needs_build():
        try:
            _build.build()
        except Exception as e:  # no nvcc on this host and no prebuilt library
            if not os.path.exists(so):
                raise ImportError(f"libradian_b200.so is missing and could not be built: {e}") from e
    lib = ctypes.CDLL(so)
    lib.radian_last_error.restype = c_char_p
    lib.radian_version.restype = c_char_p
    lib.radian_device_count.restype = c_int
    lib.radian_trim_memory.restype = c_int
    lib.radian_trim_memory.argtypes = [c_int]
    lib.radian_table_create.restype = c_int
    lib.radian_table_create.argtypes = [c_void_p, c_int, c_int, POINTER(c_void_p)]
    lib.radian_table_create_sparse.restype = c_int
    lib.radian_table_create_sparse.argtypes = [c_void_p, c_void_p, c_int, c_int, POINTER(c_void_p)]
    lib.radian_table_destroy.restype = c_int
    lib.radian_table_destroy.argtypes = [c_void_p]
    lib.radian_table_context_len.restype = c_int
    lib.radian_table_context_len.argtypes = [c_void_p]
    lib.radian_table_entropies.restype = c_int
    lib.radian_table_entropies.argtypes = [c_void_p, c_void_p]
    lib.radian_decode_workspace_bytes.restype = c_size_t
    lib.radian_decode_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int64, c_int64]
    lib.radian_decode_batch_dev.restype = c_int
    lib.radian_decode_batch_dev.argtypes = [
        c_void_p, c_int, c_void_p, c_int, c_void_p, c_int64, c_int64, c_int, c_void_p, c_int, c_double, c_double,
        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_size_t, c_void_p]
    lib.radian_decode_batch_host.restype = c_int
    lib.radian_decode_batch_host.argtypes = [
        c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_double, c_double,
        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int]
    lib.radian_decode_batch_host_reads.restype = c_int
    lib.radian_decode_batch_host_reads.argtypes = [
        c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_double, c_double,
        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int]
    lib.radian_assemble_plan.restype = c_int
    lib.radian_assemble_plan.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p, POINTER(c_int), POINTER(c_int32)]
    lib.radian_assemble_batch_dev.restype = c_int
    lib.radian_assemble_batch_dev.argtypes = [
        c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int32, c_int64, c_void_p, c_int, c_void_p]
    lib.radian_assemble_batch_host.restype = c_int
    lib.radian_assemble_batch_host.argtypes = [
        c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int]
    lib.radian_stitch_batch_host.restype = c_int
    lib.radian_stitch_batch_host.argtypes = [
        c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int]
    lib.radian_stitch_workspace_bytes.restype = c_size_t
    lib.radian_stitch_workspace_bytes.argtypes = [c_int64, c_int64, c_int64]
    lib.radian_stitch_batch_dev.restype = c_int
    lib.radian_stitch_batch_dev.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p,
                                            c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                            c_size_t, c_void_p]
    lib.radian_normalise_batch_host.restype = c_int
    lib.radian_normalise_batch_host.argtypes = [c_void_p, c_void_p, c_int, c_double, c_int, c_void_p, c_void_p,
                                                c_void_p, c_int]
    lib.radian_normalise_batch_dev.restype = c_int
    lib.radian_normalise_batch_dev.argtypes = [c_void_p, c_void_p, c_int, c_double, c_int, c_void_p, c_void_p,
                                               c_void_p, c_void_p]
    lib.radian_windows_batch_dev.restype = c_int
    lib.radian_windows_batch_dev.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]
    lib.radian_windows_plan.restype = c_int
    lib.radian_windows_plan.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]
    lib.radian_windows_batch_host.restype = c_int
    lib.radian_windows_batch_host.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int]
    lib.radian_fasta_records_host.restype = c_int
    lib.radian_fasta_records_host.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_char_p, c_void_p,
                                              c_void_p]
    return lib


lib = _load()


class RadianError(RuntimeError):
    pass


def last_error() -> str:
    return lib.radian_last_error().decode("utf-8", "replace")


def check(rc: int):
    """Map a status code to the exception class the reference would have raised."""
    if rc == OK:
        return
    msg = last_error()
    if rc == E_ARG:
        raise ValueError(msg)
    if rc == E_CONTEXT:
        raise KeyError(msg)
    if rc == E_GAP:
        raise IndexError(msg)
    raise RadianError(f"[{rc}] {msg}")


def np_ptr(a):
    return None if a is None else a.ctypes.data_as(c_void_p)
