"""Pins oracle/radian_oracle.c against outputs of the unmodified reference (tests/golden)."""
import numpy as np
import pytest

import golden_io
from oracle import oracle

FILES = ["decode_kat.npz", "decode_random.npz", "decode_synth.npz", "decode_long.npz", "decode_wide.npz",
         "decode_headline.npz"]
CASES = [c for f in FILES for c in golden_io.decode_cases(f)]


def check_scores(got, want, rel=1e-12):
    assert len(got) == len(want)
    for g, w in zip(got, want):
        if np.isinf(w):
            assert g == w
        else:
            assert abs(g - w) <= rel * max(1.0, abs(w)), (g, w)


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.name)
def test_decode_matches_reference(case):
    tab = golden_io.table(case.L, case.tseed) if case.L else None
    seq, scores, nfin, (nl, nc) = oracle.beam_search(case.mat, case.bw, tab, case.L, case.s_thr,
                                                     case.r_thr, topk=8)
    assert seq.tolist() == case.seq.tolist()
    check_scores(scores, case.scores)
    assert nl == case.n_lookup
    assert nc == case.n_combine


def test_assembly_matches_reference():
    cases = golden_io.assembly_cases()
    assert len(cases) >= 100
    n64 = 0
    for S, mats, ref in cases:
        got = oracle.assemble(mats, S)
        assert got.dtype == ref.dtype
        assert got.shape == ref.shape
        assert np.array_equal(got, ref)
        n64 += ref.dtype == np.float64
    assert n64 > 10


def test_kat_answers():
    """SURVEY.md section 4 known answers, as literal strings."""
    kat = golden_io.decode_cases("decode_kat.npz")
    assert kat[0].seq.tolist() == []              # all blank -> ''
    assert kat[1].seq.tolist() == [1, 1]          # 'CC'
    assert kat[2].seq.tolist() == [0]             # uniform, bw=3, T=3 -> 'A'


def test_stitch_matches_reference():
    """The difflib / simple_assembly restatement against both stitching fixture sets recorded from
    the unmodified reference (sequence_assembly.py:19-48, 90-97)."""
    import json
    import os

    n = n_err = 0
    for fn in ("sequence_assembly.json", "sequence_assembly_ext.json"):
        for c in json.load(open(os.path.join(golden_io.GOLDEN, fn))):
            n += 1
            if "error" in c:
                n_err += 1
                with pytest.raises(IndexError):
                    oracle.stitch(c["fragments"])
                continue
            cons, votes = oracle.stitch(c["fragments"])
            assert cons == c["consensus"]
            if "votes" in c:
                want = np.array(c["votes"], dtype=np.int64).reshape(4, -1)
                assert votes.shape == want.shape and np.array_equal(votes, want)
            else:
                assert list(votes.shape) == c["votes_shape"] and float(votes.sum()) == c["votes_sum"]
    assert n >= 130 and n_err >= 1


def test_preprocess_matches_reference():
    """mad_normalise / get_windows restatement (preprocess.py:4-49) against the reference's output
    on the five real signals of its bundled fast5 file and on synthetic edge cases: bit-exact,
    including the int64 result of np.vectorize when the first sample is clipped."""
    cases = golden_io.preprocess_cases()
    assert len(cases) >= 35
    n_err = n_int = 0
    for c in cases:
        if c["error"]:
            n_err += 1
            with pytest.raises(ValueError, match="empty" if c["error"] == 1 else "MAD is zero"):
                oracle.mad_normalise(c["signal"], c["outlier"])
            continue
        r = oracle.mad_normalise(c["signal"], c["outlier"])
        assert r.dtype == c["result"].dtype
        n_int += r.dtype == np.int64
        assert np.array_equal(r.view(np.int64), c["result"].view(np.int64))
        for W, S, nw, pad, s0, s1 in c["windows"]:
            w, p = oracle.get_windows(r, int(W), int(S))
            assert w.shape == (int(nw), int(W)) and p == int(pad)
            assert golden_io.window_checksum(w) == (s0, s1)
    assert n_err == 4 and n_int == 3


@pytest.mark.reference
def test_persistent_tie_reads_match_reference():
    """The reads of tests/test_gpu_headline.py::test_persistent_exact_ties (two labelings with bit-equal
    scores for the whole read): the oracle orders them as the unmodified reference does (stable sort
    over dict insertion order, decode.py:35-39).  Build container only."""
    from oracle import ref_loader

    dec, _, _ = ref_loader.load()
    rng_T = 300
    for seed in range(3):
        for bw in (2, 6, 16, 40):
            rng = np.random.default_rng(seed)
            p = np.zeros((rng_T, 5), np.float32 if seed % 2 else np.float64)
            p[:, 4] = 0.9
            p[:, :4] = 0.025
            p[5] = [0.3, 0.3, 0.0, 0.0, 0.4]
            t = 12
            while t < rng_T - 3:
                c = rng.integers(0, 4)
                p[t] = 0.01
                p[t, c] = 0.95
                p[t, 4] = 0.02
                t += rng.integers(4, 12)
            want = dec.beam_search(p, "ACGT", bw, None, None, None, None, None)
            got = "".join("ACGT"[s] for s in oracle.beam_search(p, bw)[0])
            assert got == want
