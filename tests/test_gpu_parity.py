"""Parity of the CUDA path (through the C ABI) with the reference: golden vectors recorded from
the unmodified reference, and the pinned C oracle on seeded synthetic reads.

Bar: decoded sequences identical; best-beam log score within 1e-9 relative (the north star
allows 1e-5; the float64 linear-domain kernel is far inside it)."""
import numpy as np
import pytest

import golden_io

pytestmark = pytest.mark.gpu

SCORE_RTOL = 1e-9
FILES = ["decode_kat.npz", "decode_random.npz", "decode_synth.npz", "decode_long.npz", "decode_wide.npz",
         "decode_headline.npz"]
CASES = [c for f in FILES for c in golden_io.decode_cases(f)]


def close(g, w, rtol=SCORE_RTOL):
    if np.isinf(w) or np.isnan(w):
        return (np.isnan(g) and np.isnan(w)) or g == w
    return abs(g - w) <= rtol * max(1.0, abs(w))


@pytest.fixture(scope="module")
def tables():
    from radian_b200.decode import RnaTable

    cache = {}

    def get(L, seed):
        if (L, seed) not in cache:
            if len(cache) > 6:
                cache.clear()
            cache[(L, seed)] = RnaTable(golden_io.table(L, seed))
        return cache[(L, seed)]

    return get


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.name)
def test_golden_decode(case, tables):
    from radian_b200 import decode

    lm = tables(case.L, case.tseed) if case.L else None
    seqs, scores, cnt = decode.beam_search_batch([case.mat], case.bw, lm, case.s_thr, case.r_thr, case.L,
                                                 return_details=True)
    want = "".join("ACGT"[s] for s in case.seq)
    assert seqs[0] == want
    assert close(scores[0, 0], case.scores[0])
    if len(case.scores) > 1 and case.bw > 1:
        assert close(scores[0, 1], case.scores[1])
    assert int(cnt[0, 0]) == case.n_lookup
    assert int(cnt[0, 1]) == case.n_combine


def test_beam_width_limit():
    from radian_b200 import decode

    with pytest.raises(ValueError):
        decode.beam_search(np.full((3, 5), 0.2, np.float32), "ACGT", 129, None, None, None, None, None)
    with pytest.raises(ValueError):
        decode.beam_search(np.full((3, 5), 0.2, np.float32), "ACGT", 0, None, None, None, None, None)


def test_dropin_signature_and_lm_quirks():
    """decode.py:100-109 signature; {} and None switch the model off (decode.py:157,180);
    the string "None" of basecall.py:48-49 is a TypeError; a wrong context length a KeyError."""
    from radian_b200 import decode

    m = np.full((3, 5), 0.2, np.float32)
    assert decode.beam_search(m, "ACGT", 3, None, None, None, None, None) == "A"
    assert decode.beam_search(m, "ACGT", 3, {}, 0.5, 0.5, 3, {}) == "A"
    assert decode.beam_search(m, "ACGU", 3, None, None, None, None, None) == "A"
    assert decode.beam_search(np.zeros((0, 5), np.float32), "ACGT", 3, None, None, None, None, None) == ""
    with pytest.raises(TypeError):
        decode.beam_search(m, "ACGT", 3, "None", 0.5, 0.5, 2, {})
    # a model that lacks contexts fails only when the search reaches one (decode.py:83): three
    # uniform frames never keep a two-symbol labeling, nine do (answers recorded from the reference)
    assert decode.beam_search(m, "ACGT", 3, {(0, 1): [0.25] * 4}, 0.5, 0.5, 2, {}) == "A"
    with pytest.raises(KeyError, match=r"\(1, 0\)"):
        decode.beam_search(np.full((9, 5), 0.2, np.float32), "ACGT", 3, {(0, 1): [0.25] * 4}, 0.5, 0.5, 2, {})
    tab = decode.RnaTable(golden_io.table(2, 1))
    with pytest.raises(KeyError):
        decode.beam_search(m, "ACGT", 3, tab, 0.5, 0.5, 3, {})


def test_sparse_model_keyerror_like_reference():
    """A dict that lacks contexts: KeyError exactly when the reference's search reaches a missing one
    (decode.py:83, recorded from the reference in decode_sparse.npz), the same result otherwise."""
    import os

    from radian_b200 import decode, synth

    z = np.load(os.path.join(golden_io.GOLDEN, "decode_sparse.npz"))
    po = so = mo = 0
    n_err = 0
    for T, L, bw, tseed, err, nseq, is64 in z["meta"]:
        mat = z["post"][po:po + T].astype(np.float64 if is64 else np.float32)
        po += T
        want = "".join("ACGT"[c] for c in z["seq"][so:so + nseq])
        so += nseq
        present = z["present"][mo:mo + 4 ** L]
        mo += 4 ** L
        tab = synth.make_table(int(L), int(tseed))
        lm = {tuple((int(i) >> (2 * (int(L) - 1 - j))) & 3 for j in range(int(L))): tab[i].tolist()
              for i in np.flatnonzero(present)}
        if err:
            n_err += 1
            with pytest.raises(KeyError):
                decode.beam_search(mat, "ACGT", int(bw), lm, 0.5, 0.5, int(L), {})
        else:
            assert decode.beam_search(mat, "ACGT", int(bw), lm, 0.5, 0.5, int(L), {}) == want
    assert 10 < n_err < 35


def test_dict_lm_equals_dense_table():
    from radian_b200 import decode
    from radian_b200 import synth

    dense = synth.make_table(3, 9)
    lm = {}
    for i in range(64):
        lm[((i >> 4) & 3, (i >> 2) & 3, i & 3)] = dense[i].tolist()
    post, off = synth.make_reads(np.array([40]), seed=5)
    m = post.numpy().astype(np.float64)
    a = decode.beam_search(m, "ACGT", 6, lm, 0.5, 0.5, 3, {})
    b = decode.beam_search(m, "ACGT", 6, decode.RnaTable(dense), 0.5, 0.5, 3, None)
    assert a == b and len(a) > 10


def test_table_entropies_match_reference_formula():
    """decode.py:73-76 on the host in float64: bit-exact against the same formula in Python."""
    import math

    from radian_b200 import decode

    dense = golden_io.table(4, 3)
    dense[5] = [0.5, 0.5, 0.0, 0.0]
    dense[6] = [1.0, 0.0, 0.0, 0.0]
    got = decode.RnaTable(dense).entropies()
    # iterate numpy scalars exactly like decode.py:75-76 (CPython >= 3.12 sums exact Python
    # floats with compensation, np.float64 scalars go through the plain left-to-right path)
    want = np.array([-sum([p * math.log(p) for p in row[row > 0]]) for row in dense])
    assert np.array_equal(got, want)


@pytest.mark.parametrize("bw,L,f64", [(6, 0, False), (16, 6, True), (16, 11, True), (8, 4, False),
                                      (32, 5, True), (1, 3, True), (2, 0, True), (17, 2, False),
                                      (33, 0, False), (64, 6, True), (64, 11, False), (48, 3, True),
                                      (100, 4, False), (128, 0, True), (65, 5, True),
                                      (128, 7, False), (128, 5, True), (96, 9, False)])
def test_batch_vs_oracle(bw, L, f64):
    """A mixed-length batch through one launch against the pinned C oracle, read by read."""
    from oracle import oracle
    from radian_b200 import decode, synth

    nb = np.array([5, 80, 33, 150, 1, 60, 240, 18, 99, 47, 12, 130])
    post, off = synth.make_reads(nb, seed=100 + bw + L)
    post = post.numpy()
    off = off.numpy()
    if f64:
        post = post.astype(np.float64)
    mats = [post[off[i]:off[i + 1]] for i in range(len(nb))]
    mats.append(post[:0])  # an empty read
    tab = synth.make_table(L, 21) if L else None
    lm = decode.RnaTable(tab) if L else None
    seqs, scores, cnt = decode.beam_search_batch(mats, bw, lm, 0.5, 0.5, L, return_details=True)
    for i, m in enumerate(mats):
        oseq, osc, _, (nl, nc) = oracle.beam_search(m, bw, tab, L, 0.5, 0.5, topk=2)
        assert seqs[i] == "".join("ACGT"[s] for s in oseq), f"read {i}"
        assert close(scores[i, 0], osc[0]), f"read {i}"
        assert int(cnt[i, 0]) == nl and int(cnt[i, 1]) == nc


@pytest.mark.parametrize("bw,nb", [(16, (3000, 2500)), (64, (1500, 900)), (128, (1200, 700))])
def test_arena_compaction_long_read(bw, nb):
    """A read long enough to fill the back-pointer arena many times (compaction + flush)."""
    from oracle import oracle
    from radian_b200 import decode, synth

    post, off = synth.make_reads(np.array(nb), seed=77)
    post = post.numpy()
    off = off.numpy()
    mats = [post[off[i]:off[i + 1]] for i in range(2)]
    tab = synth.make_table(7, 2)
    seqs, scores, _ = decode.beam_search_batch(mats, bw, decode.RnaTable(tab), 0.5, 0.5, 7, return_details=True)
    for i, m in enumerate(mats):
        oseq, osc, _, _ = oracle.beam_search(m, bw, tab, 7, 0.5, 0.5, topk=1)
        assert seqs[i] == "".join("ACGT"[s] for s in oseq)
        assert close(scores[i, 0], osc[0])


def test_streamed_host_batch_order_and_results():
    """The _host entry point queues reads longest first and streams them while the kernel runs;
    results must come back in the caller's order, identical to one-read-at-a-time calls."""
    from radian_b200 import decode, synth

    nb = synth.read_lengths(300, 11, median=150, lo=1, hi=2500)
    post, off = synth.make_reads(nb, seed=12)
    post = post.numpy()
    off = off.numpy()
    mats = [post[off[i]:off[i + 1]] for i in range(len(nb))]
    mats.insert(7, post[:0])
    tab = decode.RnaTable(synth.make_table(5, 4))
    seqs, scores, cnt = decode.beam_search_batch(mats, 16, tab, 0.5, 0.5, 5, return_details=True)
    for i in (0, 7, 8, 100, 299, 300):
        one, sc1, c1 = decode.beam_search_batch([mats[i]], 16, tab, 0.5, 0.5, 5, return_details=True)
        assert one[0] == seqs[i]
        assert sc1[0, 0] == scores[i, 0]
        assert (c1[0][:3] == cnt[i][:3]).all()  # (the fourth is a diagnostic that depends on the warp-mate)
    assert seqs[7] == ""


def test_assembly_golden():
    from radian_b200 import matrix_assembly

    cases = golden_io.assembly_cases()
    for S, mats, ref in cases:
        got = matrix_assembly.assemble_matrices(mats, S)
        assert got.dtype == ref.dtype
        assert got.shape == ref.shape
        assert np.array_equal(got, ref)


def test_assembly_batch_and_errors():
    from radian_b200 import matrix_assembly, synth

    post, off = synth.make_reads(np.array([70, 20, 130]), seed=3)
    post = post.numpy()
    off = off.numpy()
    batch = [synth.split_windows(post[off[i]:off[i + 1]], 1024, 128) for i in range(3)]
    outs = matrix_assembly.assemble_batch(batch, 128)
    from oracle import oracle

    for mats, got in zip(batch, outs):
        want = oracle.assemble(mats, 128)
        assert got.dtype == want.dtype and np.array_equal(got, want)
    assert matrix_assembly.assemble_matrices([], 128).shape == (0,)
    with pytest.raises(IndexError):  # create_vstack's list index error on a gap
        matrix_assembly.assemble_matrices([post[:3], post[:3]], 10)
    with pytest.raises(ValueError):
        matrix_assembly.assemble_matrices([post[:3]], 0)


def test_global_pipeline_matches_reference_shape():
    """config 1 shape end to end: windows -> assemble -> global decode, vs the oracle."""
    from oracle import oracle
    from radian_b200 import decode, matrix_assembly, synth

    post, off = synth.make_reads(np.array([110]), seed=8)
    post = post.numpy()
    mats = synth.split_windows(post, 1024, 128)
    mat = matrix_assembly.assemble_matrices(mats, 128)
    assert mat.dtype == np.float64
    tab = synth.make_table(11, 5)
    got = decode.beam_search(mat, "ACGT", 6, decode.RnaTable(tab), 0.5, 0.5, 11, {})
    oseq, _, _, _ = oracle.beam_search(oracle.assemble(mats, 128), 6, tab, 11, 0.5, 0.5)
    assert got == "".join("ACGT"[s] for s in oseq)


def test_device_resident_path():
    """decode_batch_device on torch CUDA tensors == host path."""
    import torch

    from radian_b200 import decode, synth

    nb = synth.read_lengths(24, 4, median=120, lo=20, hi=400)
    post, off = synth.make_reads(nb, seed=6, device="cuda")
    tab = synth.make_table(6, 1)
    lm = decode.RnaTable(tab)
    T = (off[1:] - off[:-1])
    order = torch.argsort(T, descending=True).to(torch.int32)
    res = decode.decode_batch_device(post, off, 16, lm, 0.5, 0.5, max_frames=int(T.max()), order=order, counters=True)
    torch.cuda.synchronize()
    assert int(res.status.abs().sum()) == 0
    dev = res.strings()
    p = post.cpu().numpy()
    o = off.cpu().numpy()
    host = decode.beam_search_batch([p[o[i]:o[i + 1]] for i in range(len(nb))], 16, lm, 0.5, 0.5, 6)
    assert dev == host


def test_stitch_golden():
    """Chunk-mode stitching (sequence_assembly.py:19-48, 90-97) against outputs of the reference:
    votes bit-exact, consensus identical, the reference's IndexError reproduced."""
    import json
    import os

    from radian_b200.sequence_assembly import index2base, simple_assembly, stitch_batch

    batch, want = [], []
    for fn in ("sequence_assembly.json", "sequence_assembly_ext.json"):
        for c in json.load(open(os.path.join(golden_io.GOLDEN, fn))):
            if "error" in c:
                with pytest.raises(IndexError):
                    simple_assembly(c["fragments"])
                continue
            votes = simple_assembly(c["fragments"])
            assert votes.dtype == np.float64
            if "votes" in c:
                ref = np.array(c["votes"], dtype=np.float64).reshape(4, -1)
                assert votes.shape == ref.shape and np.array_equal(votes, ref)
            else:
                assert list(votes.shape) == c["votes_shape"] and float(votes.sum()) == c["votes_sum"]
            assert index2base(np.argmax(votes, axis=0)) == c["consensus"]
            batch.append(c["fragments"])
            want.append(c["consensus"])
    assert len(batch) > 130
    assert stitch_batch(batch) == want  # all reads in one call
    with pytest.raises(KeyError):
        simple_assembly(["ACGT", "ACNT"])


def test_chunk_mode_pipeline_vs_oracle():
    """config 4(i): per-chunk decode with the model off, then stitching, vs the oracle end to end."""
    import types

    from oracle import oracle
    from radian_b200 import basecall, synth

    post, off = synth.make_reads(np.array([90, 40, 3]), seed=21)
    post = post.numpy()
    off = off.numpy()
    chunk_lists = [synth.split_windows(post[off[i]:off[i + 1]], 256, 32) for i in range(3)]
    args = types.SimpleNamespace(decode_type="chunk", beam_width=6, step_size=32)
    got = basecall.basecall_batch(["a", "b", "c"], chunk_lists, args, None)
    for mats, g in zip(chunk_lists, got):
        frags = ["".join("ACGT"[s] for s in oracle.beam_search(m, 6)[0]) for m in mats]
        assert g == oracle.stitch(frags)[0]


def test_preprocess_golden():
    """mad_normalise / get_windows (preprocess.py:4-49) against the reference's outputs on the five
    real signals of its bundled fast5 file and on synthetic edge cases: bit-exact, same dtypes
    (incl. np.vectorize's int64 result), same ValueErrors; one batched call for all reads."""
    from radian_b200 import preprocess

    cases = golden_io.preprocess_cases()
    n_int = 0
    for c in cases:
        if c["error"]:
            with pytest.raises(ValueError, match="empty" if c["error"] == 1 else "MAD is zero"):
                preprocess.mad_normalise(c["signal"], c["outlier"])
            continue
        r = preprocess.mad_normalise(c["signal"], c["outlier"])
        assert r.dtype == c["result"].dtype
        n_int += r.dtype == np.int64
        assert np.array_equal(r.view(np.int64), c["result"].view(np.int64))
        for W, S, nw, pad, s0, s1 in c["windows"]:
            w, p = preprocess.get_windows(r, int(W), int(S))
            assert w.dtype == r.dtype and w.shape == (int(nw), int(W)) and p == int(pad)
            assert golden_io.window_checksum(w) == (s0, s1)
    assert n_int == 3
    ints = [c for c in cases if isinstance(c["outlier"], int) and c["outlier"] == 4]
    res = preprocess.mad_normalise_batch([c["signal"] for c in ints], 4)
    for c, r in zip(ints, res):
        if c["error"]:
            assert isinstance(r, ValueError)
        else:
            assert r.dtype == c["result"].dtype and np.array_equal(r.view(np.int64), c["result"].view(np.int64))
    with pytest.raises(ValueError, match="Step size must be > 0"):
        preprocess.get_windows(np.zeros(10), 4, 0)
    with pytest.raises(ValueError, match="<= window size"):
        preprocess.get_windows(np.zeros(10), 4, 5)
    w, p = preprocess.get_windows(np.zeros(0), 4, 2)
    assert w.shape == (1, 4) and p == 4


def test_preprocess_large_random_vs_oracle():
    """A long read and many short ones against the pinned oracle (medians over a wide value range)."""
    from oracle import oracle
    from radian_b200 import preprocess

    rng = np.random.default_rng(5)
    sigs = [np.clip(np.rint(rng.normal(650, 80, 300000)), -32768, 32767).astype(np.int16)]
    sigs += [rng.integers(-2000, 3000, int(n)).astype(np.int16) for n in rng.integers(2, 5000, 64)]
    for s, r in zip(sigs, preprocess.mad_normalise_batch(sigs, 4)):
        want = oracle.mad_normalise(s, 4)
        assert r.dtype == want.dtype and np.array_equal(r.view(np.int64), want.view(np.int64))


def _toy_sig_model(windows):
    """Deterministic stand-in for the out-of-scope signal model: (n, W) z-scores -> (n, W, 5)
    float32 posteriors, blank last."""
    z = np.asarray(windows, dtype=np.float64)
    lg = np.stack([4 * np.sin(3 * z), 4 * np.cos(2 * z), 4 * np.sin(5 * z + 1), 4 * np.cos(7 * z),
                   5.5 - np.abs(z)], axis=-1)
    e = np.exp(lg - lg.max(-1, keepdims=True))
    return (e / e.sum(-1, keepdims=True)).astype(np.float32)


@pytest.mark.parametrize("mode", ["global", "chunk"])
def test_cli_from_raw_fast5_vs_oracle(tmp_path, mode):
    """config 1 shape from the raw signal of the reference's own fast5 file: read, normalise,
    window, (toy) signal model, assemble + decode or per-chunk decode + stitch, FASTA; against the
    oracle doing the same steps (basecall.py:70-141)."""
    import shutil

    from oracle import oracle
    from radian_b200 import basecall, fast5

    indir = tmp_path / "in"
    outdir = tmp_path / "out"
    indir.mkdir()
    outdir.mkdir()
    src = __import__("os").path.join(golden_io.GOLDEN, "reads.fast5")
    shutil.copy(src, indir / "reads.fast5")
    basecall.main([str(indir), str(outdir), "--rna-model", "None", "--decode-type", mode, "--beam-width", "6",
                   "--chunk-len", "512", "--step-size", "64"], sig_model=_toy_sig_model)
    got = (outdir / "reads-0.fasta").read_text().split("\n")
    want = []
    for rid, sig in fast5.reads(src):
        norm = oracle.mad_normalise(sig, 4)
        win, pad = oracle.get_windows(norm, 512, 64)
        mats = list(_toy_sig_model(win))
        mats[-1] = mats[-1][:-pad]
        if mode == "global":
            seq = "".join("ACGT"[s] for s in oracle.beam_search(oracle.assemble(mats, 64), 6)[0])
        else:
            seq = oracle.stitch(["".join("ACGT"[s] for s in oracle.beam_search(m, 6)[0]) for m in mats])[0]
        want += [f">{rid}", seq[::-1]]
    assert got[:len(want)] == want and len(want) == 10


def test_near_tie_counter():
    """out_counters[2]: frames whose selection ranked two candidates within 2^-40 of each other.
    Posteriors drawn from a handful of values are full of them; realistic posteriors have almost
    none (a beam search on those never hangs on rounding noise)."""
    from radian_b200 import decode, synth

    rng = np.random.default_rng(2)
    lg = np.round(rng.normal(0, 1.5, (300, 5)))
    tie = np.exp(lg)
    tie = (tie / tie.sum(1, keepdims=True)).astype(np.float32)
    post, off = synth.make_reads(np.array([40, 60, 25]), seed=9)
    post = post.numpy()
    off = off.numpy()
    mats = [tie] + [post[off[i]:off[i + 1]] for i in range(3)]
    for bw in (6, 64):
        _, scores, cnt = decode.beam_search_batch(mats, bw, None, None, None, None, return_details=True)
        assert cnt.shape == (4, 4) and int(cnt[0, 2]) > 20
        assert int(cnt[1:, 2].sum()) <= 3


def test_stitch_device_resident_matches_host():
    """radian_stitch_batch_dev on the decoder's own output slots == the host entry point."""
    import torch

    from radian_b200 import decode, sequence_assembly, synth

    post, off = synth.make_reads(np.array([70, 30, 45]), seed=33, device="cuda")
    p = post.cpu().numpy()
    o = off.cpu().numpy()
    chunk_lists = [synth.split_windows(p[o[i]:o[i + 1]], 256, 32) for i in range(3)]
    flat = [m for cl in chunk_lists for m in cl]
    rfr = np.zeros(4, np.int64)
    rfr[1:] = np.cumsum([len(cl) for cl in chunk_lists])
    cro = np.zeros(len(flat) + 1, np.int64)
    cro[1:] = np.cumsum([len(m) for m in flat])
    chunks = torch.from_numpy(np.concatenate(flat)).cuda()
    d_cro = torch.from_numpy(cro).cuda()
    res = decode.decode_batch_device(chunks, d_cro, 6, None, max_frames=256)
    seq, ooff, ln, st = sequence_assembly.stitch_device(res.seq, res.seq_offsets[:-1].contiguous(), res.lengths,
                                                        torch.from_numpy(rfr).cuda())
    torch.cuda.synchronize()
    assert int(st.abs().sum()) == 0
    hs, ho, hl = seq.cpu().numpy(), ooff.cpu().numpy(), ln.cpu().numpy()
    got = ["".join("ACGT"[s] for s in hs[ho[r]:ho[r] + hl[r]]) for r in range(3)]
    frags = decode.beam_search_batch(flat, 6, None, None, None, None)
    want = sequence_assembly.stitch_batch([frags[rfr[r]:rfr[r + 1]] for r in range(3)])
    assert got == want and all(len(g) > 20 for g in got)


def test_host_entry_point_is_reentrant():
    """Several host threads decoding at once through the same table handle (one call each, own
    streams and scratch) get the results of a single-threaded run."""
    from concurrent.futures import ThreadPoolExecutor

    from radian_b200 import decode, synth

    tab = decode.RnaTable(synth.make_table(6, 3))
    batches = []
    for k in range(6):
        post, off = synth.make_reads(synth.read_lengths(20, 50 + k, median=60, lo=5, hi=200), seed=70 + k)
        p, o = post.numpy(), off.numpy()
        batches.append([p[o[i]:o[i + 1]] for i in range(20)])
    want = [decode.beam_search_batch(b, 16, tab, 0.5, 0.5, 6) for b in batches]
    with ThreadPoolExecutor(6) as ex:
        got = list(ex.map(lambda b: decode.beam_search_batch(b, 16, tab, 0.5, 0.5, 6), batches * 3))
    assert got == want * 3
    decode.trim_memory()  # hands the pooled device buffers back; the next call allocates afresh
    assert decode.beam_search_batch(batches[0], 16, tab, 0.5, 0.5, 6) == want[0]


@pytest.mark.parametrize("bw", [16, 64])
def test_stalled_transfer_is_survived(bw, monkeypatch):
    """If the copies of a streamed call stand still (here: held back for 1.5 s by a test hook, in
    real life e.g. by another thread's cudaFree waiting for the device), the kernel stops waiting
    after half a second and the reads it did not get are decoded by a second launch: same
    results, no hang."""
    from radian_b200 import decode, synth

    post, off = synth.make_reads(synth.read_lengths(40, 8, median=80, lo=5, hi=300), seed=44)
    p, o = post.numpy(), off.numpy()
    mats = [p[o[i]:o[i + 1]] for i in range(40)]
    tab = decode.RnaTable(synth.make_table(5, 2))
    want = decode.beam_search_batch(mats, bw, tab, 0.5, 0.5, 5, return_details=True)
    monkeypatch.setenv("RADIAN_TEST_STALL_MS", "1500")
    got = decode.beam_search_batch(mats, bw, tab, 0.5, 0.5, 5, return_details=True)
    assert got[0] == want[0] and np.array_equal(got[1], want[1], equal_nan=True) and np.array_equal(got[2][:, :3], want[2][:, :3])
