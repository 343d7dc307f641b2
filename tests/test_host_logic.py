"""Host-side logic that needs no GPU: assembly planning through the C ABI, sharding, the CLI
surface, the synthetic generators and the chunk-mode stitcher."""
import ctypes
import hashlib
import json
import os

import numpy as np
import pytest

import golden_io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def plan(chunk_lens_per_read, step):
    from radian_b200 import _native

    cro = [0]
    rcr = [0]
    for lens in chunk_lens_per_read:
        for ln in lens:
            cro.append(cro[-1] + ln)
        rcr.append(len(cro) - 1)
    cro = np.asarray(cro, np.int64)
    rcr = np.asarray(rcr, np.int64)
    n = len(chunk_lens_per_read)
    rows = np.zeros(n, np.int64)
    ov = ctypes.c_int(0)
    mx = ctypes.c_int32(0)
    rc = _native.lib.radian_assemble_plan(_native.np_ptr(cro), _native.np_ptr(rcr), n, step, _native.np_ptr(rows),
                                          ctypes.byref(ov), ctypes.byref(mx))
    _native.check(rc)
    return rows.tolist(), bool(ov.value), int(mx.value)


def test_assemble_plan_matches_golden_shapes():
    """Row counts and float64 promotion of every golden assembly case (matrix_assembly.py:12-44)."""
    for S, mats, ref in golden_io.assembly_cases():
        rows, ov, mx = plan([[len(m) for m in mats]], S)
        assert rows == [ref.shape[0]]
        assert ov == (ref.dtype == np.float64)
        assert abs(mx) == max(len(m) for m in mats)


def test_assemble_plan_closed_form_and_errors():
    # reference windowing: T = (n-1)*S + len(last)  
    rows, ov, mx = plan([[1024] * 9 + [1000], [700]], 128)
    assert rows == [9 * 128 + 1000, 700] and ov and mx == 1024
    rows, ov, _ = plan([[10, 10, 10]], 10)       # step == window: no overlap -> float32 result
    assert rows == [30] and not ov
    with pytest.raises(IndexError):               # create_vstack cannot leave a gap
        plan([[3, 3]], 10)
    with pytest.raises(ValueError):
        plan([[3]], 0)


def test_lpt_shards_are_balanced_and_complete():
    from radian_b200 import parallel, synth

    fc = synth.read_lengths(1000, 1) * 43
    for world in (1, 2, 8):
        shards = parallel.lpt_shards(fc, world)
        allidx = np.sort(np.concatenate(shards))
        assert np.array_equal(allidx, np.arange(1000))
        loads = np.array([fc[s].sum() for s in shards])
        assert loads.max() - loads.min() <= fc.max()
        for s in shards:
            assert np.all(np.diff(fc[s]) <= 0)   # longest first within a rank


def test_cli_flags_match_reference():
    """Names, defaults and types of basecall.py:21-35."""
    from radian_b200 import basecall

    a = basecall.build_parser().parse_args(["in", "out"])
    assert (a.chunk_len, a.step_size, a.batch_size, a.outlier_clip) == (1024, 128, 32, 4)
    assert (a.beam_width, a.decode_type, a.sig_threshold, a.rna_threshold, a.context_len) == (6, "global", 0.5, 0.5, 11)
    assert a.rna_model == "models/rnamodel_12mer_pc.json"
    b = basecall.build_parser().parse_args(["in", "out", "--beam-width", "16", "--decode-type", "chunk",
                                           "--sig-threshold", "0.3", "--rna-threshold", "0.9", "--context-len", "12",
                                           "--chunk-len", "512", "--step-size", "64"])
    assert (b.beam_width, b.decode_type, b.sig_threshold, b.rna_threshold, b.context_len, b.chunk_len, b.step_size) == \
        (16, "chunk", 0.3, 0.9, 12, 512, 64)
    with pytest.raises(SystemExit):
        basecall.build_parser().parse_args(["in", "out", "--decode-type", "local"])


def test_fasta_writer_rollover(tmp_path):
    """>{id}\\n{seq[::-1]}\\n records, a new file every 1000 reads (basecall.py:129-138)."""
    from radian_b200.basecall import FastaWriter

    w = FastaWriter(str(tmp_path), per_file=3)
    for i in range(7):
        w.write(f"r{i}", "ACG" + "T" * i)
    w.close()
    files = sorted(os.listdir(tmp_path))
    assert files == ["reads-0.fasta", "reads-1.fasta", "reads-2.fasta"]
    assert open(tmp_path / "reads-0.fasta").read() == ">r0\nGCA\n>r1\nTGCA\n>r2\nTTGCA\n"
    assert open(tmp_path / "reads-2.fasta").read() == ">r6\nTTTTTTGCA\n"


def test_synth_table_is_bit_reproducible():
    from radian_b200 import synth

    t = synth.make_table(5, 7)
    assert t.shape == (1024, 4)
    assert np.allclose(t.sum(1), 1.0)
    assert hashlib.sha256(t.tobytes()).hexdigest() == hashlib.sha256(synth.make_table(5, 7, chunk=100).tobytes()).hexdigest()
    frac = (synth.table_entropy(synth.make_table(8, 1)) < 0.5).mean()
    assert 0.2 < frac < 0.4


def test_synth_reads_shape():
    from radian_b200 import synth

    nb = synth.read_lengths(2000, 3)
    assert nb.min() >= 200 and nb.max() <= 10000 and 1300 < nb.mean() < 1800
    post, off = synth.make_reads(np.array([30, 50]), seed=1)
    assert post.shape[1] == 5 and off[0] == 0 and off[-1] == post.shape[0]
    p = post.numpy()
    assert (p[:, :4] == 0).any() and p[:, 4].mean() > 0.8
    mats = synth.split_windows(p[: off[1]], 1024, 128)
    assert sum(len(m) for m in mats[:1]) == min(1024, off[1])


def test_fast5_reader_on_the_reference_file():
    """radian_b200.fast5 on the reference's bundled multi-read file (tests/golden/reads.fast5 is a
    copy of radian/data/reads.fast5): ids and int16 signals as ont_fast5_api returns them
    (basecall.py:70-76); the signals are the ones the preprocess fixtures were recorded on."""
    from radian_b200 import fast5

    got = list(fast5.reads(os.path.join(golden_io.GOLDEN, "reads.fast5")))
    assert [len(s) for _, s in got] == [12833, 4863, 11388, 14799, 9905]
    assert got[0][0] == "00256416-5423-47a9-ad91-54a87a6be5e5"
    assert all(s.dtype == np.int16 for _, s in got)
    cases = golden_io.preprocess_cases()
    for (rid, sig), c in zip(got, cases[:5]):
        assert np.array_equal(sig, c["signal"])
    with pytest.raises(fast5.Fast5Error):
        list(fast5.reads(os.path.join(golden_io.GOLDEN, "assembly.npz")))


def test_windows_plan_matches_reference_counts():
    """radian_windows_plan (host only): window counts and pad_end of preprocess.get_windows
    (preprocess.py:9-20) as recorded from the reference, and its two ValueErrors."""
    from radian_b200 import _native

    lib = _native.lib
    for c in golden_io.preprocess_cases():
        n = len(c["signal"])
        for W, S, nw, pad, _, _ in c["windows"]:
            off = np.array([0, n], np.int64)
            cnt = np.zeros(1, np.int64)
            p = np.zeros(1, np.int32)
            _native.check(lib.radian_windows_plan(_native.np_ptr(off), 1, int(W), int(S), _native.np_ptr(cnt),
                                                  _native.np_ptr(p)))
            assert (int(cnt[0]), int(p[0])) == (int(nw), int(pad))
    off = np.array([0, 10], np.int64)
    cnt = np.zeros(1, np.int64)
    p = np.zeros(1, np.int32)
    with pytest.raises(ValueError, match="Step size must be > 0"):
        _native.check(lib.radian_windows_plan(_native.np_ptr(off), 1, 4, 0, _native.np_ptr(cnt), _native.np_ptr(p)))
    with pytest.raises(ValueError, match="<= window size"):
        _native.check(lib.radian_windows_plan(_native.np_ptr(off), 1, 4, 5, _native.np_ptr(cnt), _native.np_ptr(p)))


def test_new_entry_points_fail_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from radian_b200 import _native, preprocess, sequence_assembly

    with pytest.raises(_native.RadianError, match="no CPU fallback"):
        preprocess.mad_normalise(np.arange(10, dtype=np.int16), 4)
    with pytest.raises(_native.RadianError, match="no CPU fallback"):
        preprocess.get_windows(np.zeros(10), 4, 2)
    with pytest.raises(_native.RadianError, match="no CPU fallback"):
        sequence_assembly.simple_assembly(["ACGT", "CGTA"])
    with pytest.raises(TypeError):
        preprocess.mad_normalise(np.arange(10, dtype=np.float32), 4)
    with pytest.raises(KeyError):
        sequence_assembly.simple_assembly(["ACGT", "ACNT"])


def test_fasta_records_match_reference_format():
    """radian_fasta_records_host == f">{id}\\n{seq[::-1]}\\n" per read (basecall.py:129); host only."""
    from radian_b200 import fasta

    rng = np.random.default_rng(4)
    n = 700
    ln = rng.integers(0, 60, n)
    ln[5] = 0
    so = np.zeros(n + 1, np.int64)
    so[1:] = np.cumsum(ln + rng.integers(0, 9, n))  # slots larger than the sequences
    seq = rng.integers(0, 4, int(so[-1])).astype(np.uint8)
    ids = [f"read-{i:04x}" + "x" * int(i % 5) for i in range(n)]
    txt = bytes(fasta.format_records(ids, seq, so, ln))
    want = "".join(f">{ids[i]}\n{''.join('ACGT'[s] for s in seq[so[i]:so[i] + ln[i]])[::-1]}\n" for i in range(n))
    assert txt.decode("ascii") == want
    assert bytes(fasta.format_records(fasta.pack_ids(ids), seq, so, ln, bases="ACGU")).decode() == want.replace("T", "U")
    assert len(fasta.format_records([], seq[:0], np.zeros(1, np.int64), np.zeros(0, np.int64))) == 0
    with pytest.raises(ValueError):
        fasta.format_records(ids, seq, so, ln[:-1])


def test_rna_json_dense_cache(tmp_path):
    """basecall.py:47-57 JSON -> dense table: same rows as the dict, absent contexts marked, and a
    second load comes from the .npz written beside the JSON; an edited JSON is parsed again."""
    import json
    import os
    import time

    from radian_b200 import decode, synth

    L = 5
    tab = synth.make_table(L, 3)
    names = ["".join("ACGT"[(i >> (2 * (L - 1 - j))) & 3] for j in range(L)) for i in range(4 ** L)]
    p = tmp_path / "model.json"
    p.write_text(json.dumps({k: tab[i].tolist() for i, k in enumerate(names) if i % 7}))
    dense, present = decode.load_rna_json(str(p))
    idx = np.flatnonzero(present)
    assert len(idx) == 4 ** L - (4 ** L + 6) // 7 and np.array_equal(dense[idx], tab[idx])
    caches = [f for f in os.listdir(tmp_path) if f.endswith(".radian_dense.npz")]
    assert len(caches) == 1
    stamp = os.stat(tmp_path / caches[0]).st_mtime_ns
    d2, p2 = decode.load_rna_json(str(p))
    assert np.array_equal(d2, dense) and np.array_equal(p2, present)
    assert os.stat(tmp_path / caches[0]).st_mtime_ns == stamp  # read, not rewritten
    time.sleep(0.01)
    p.write_text(json.dumps({k: tab[i].tolist() for i, k in enumerate(names)}))  # now complete
    d3, p3 = decode.load_rna_json(str(p))
    assert p3.all() and np.array_equal(d3, tab)
    d4, _ = decode.load_rna_json(str(p), cache=False)
    assert np.array_equal(d4, tab)


def test_fast5_chunk_filters():
    """HDF5 filter pipeline of a chunk: gzip, byte shuffle, Fletcher-32 trailer, per-chunk skip mask;
    VBZ says so instead of returning garbage."""
    import zlib

    from radian_b200 import fast5

    x = np.arange(-700, 900, dtype="<i2")
    shuffled = x.view(np.uint8).reshape(-1, 2).T.tobytes()
    got = fast5.unfilter(zlib.compress(shuffled) + b"\0\0\0\0", [fast5.FILTER_SHUFFLE, fast5.FILTER_DEFLATE,
                                                                   fast5.FILTER_FLETCHER32], 0, 2)
    assert np.array_equal(np.frombuffer(got, "<i2"), x)
    got = fast5.unfilter(zlib.compress(x.tobytes()), [fast5.FILTER_SHUFFLE, fast5.FILTER_DEFLATE], 1, 2)  # shuffle skipped
    assert np.array_equal(np.frombuffer(got, "<i2"), x)
    with pytest.raises(NotImplementedError, match="VBZ"):
        fast5.unfilter(b"abc", [fast5.FILTER_VBZ], 0, 2)


def test_posterior_batch_files_without_pickles(tmp_path, monkeypatch):
    """The CLI's posterior batches are plain arrays; object arrays (pickles) are refused unless the
    user opts in."""
    from radian_b200 import basecall

    rng = np.random.default_rng(3)
    w = [rng.random((1024, 5), dtype=np.float32) for _ in range(3)] + [rng.random((77, 5), dtype=np.float32)]
    np.savez(tmp_path / "a.npz", r1=np.concatenate(w), r1__lens=np.array([len(m) for m in w]),
             r2=np.stack(w[:3]))
    ids, lists = basecall.load_posterior_batch(tmp_path / "a.npz")
    assert ids == ["r1", "r2"] and [len(x) for x in lists] == [4, 3]
    assert all(np.array_equal(a, b) for a, b in zip(lists[0], w))
    obj = np.empty(2, dtype=object)
    obj[0], obj[1] = w[0], w[3]
    np.savez(tmp_path / "b.npz", r3=obj)
    monkeypatch.delenv("RADIAN_ALLOW_PICKLE", raising=False)
    with pytest.raises(ValueError, match="pickle"):
        basecall.load_posterior_batch(tmp_path / "b.npz")
    monkeypatch.setenv("RADIAN_ALLOW_PICKLE", "1")
    ids, lists = basecall.load_posterior_batch(tmp_path / "b.npz")
    assert ids == ["r3"] and lists[0][1].shape == (77, 5)
