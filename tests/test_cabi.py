"""The C-ABI library loads and exports every symbol include/radian_b200.h declares; without a GPU
the compute entry points fail loudly instead of falling back to anything."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "radian_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(radian_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from radian_b200 import _native

    names = declared_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(_native.lib, n), f"{n} declared in the header but not exported"
    assert sorted(_native.EXPORTS) == names


def test_version_and_error_string():
    from radian_b200 import _native

    assert b"sm_100a" in _native.lib.radian_version()
    assert isinstance(_native.last_error(), str)


def test_oracle_not_imported_by_product():
    """Only tests/, smoke() and bench.py may touch oracle/ (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "radian_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "radian_oracle" not in src, f


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from radian_b200 import _native, decode, matrix_assembly

    assert _native.lib.radian_device_count() == 0
    m = np.full((3, 5), 0.2, np.float32)
    with pytest.raises(_native.RadianError, match="no CPU fallback"):
        decode.beam_search(m, "ACGT", 3, None, None, None, None, None)
    with pytest.raises(_native.RadianError):
        matrix_assembly.assemble_matrices([m, m], 1)
    with pytest.raises(_native.RadianError, match="no CPU fallback"):
        decode.RnaTable(np.full((4, 4), 0.25))


def test_argument_errors_need_no_gpu():
    from radian_b200 import decode

    m = np.full((3, 5), 0.2, np.float32)
    with pytest.raises(ValueError):
        decode.beam_search(m, "ACGT", 0, None, None, None, None, None)
    with pytest.raises(ValueError):
        decode.beam_search(m, "ACGT", 129, None, None, None, None, None)
    with pytest.raises(ValueError):
        decode.beam_search(m, "ACG", 3, None, None, None, None, None)
    with pytest.raises(ValueError):
        decode.beam_search(np.zeros((3, 4), np.float32), "ACGT", 3, None, None, None, None, None)
    with pytest.raises(TypeError):
        decode.beam_search(m, "ACGT", 3, "None", 0.5, 0.5, 2, {})
