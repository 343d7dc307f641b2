"""Parity on BASELINE.json's own configurations (the reference-recorded fixtures for them are in
tests/golden/decode_headline.npz and run through test_gpu_parity.py::test_golden_decode):

  C3  RNA-LM global decode, bw 16, 12-symbol context, float32 posteriors, LogNormal read lengths
      including one 10 kb read -- exactly what bench.py times;
  C4  --chunk-len 1024 --step-size 128, bw 16, both decode types;
  C5  corners of the sweep: bw 6 / 64 x context 6 / 12 x 0.5 / 10 kb;
  and float64 matrices far outside the float32 range (the reference's log-domain scores have no
  range limit; decode.py:16-17, 172-175).

Checker: the pinned C oracle (oracle/radian_oracle.c) on the same seeded inputs.  Reference
semantics matched: radian/decode.py:141-210, basecall.py:99-123."""
import os
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

THREADS = os.cpu_count() or 1


def close(g, w, rtol=1e-9):
    if np.isinf(w) or np.isnan(w):
        return (np.isnan(g) and np.isnan(w)) or g == w
    return abs(g - w) <= rtol * max(1.0, abs(w))


def split(post, off):
    return [post[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def oracle_batch(mats, bw, tab, L, s_thr=0.5, r_thr=0.5):
    from oracle import oracle

    fo = np.zeros(len(mats) + 1, dtype=np.int64)
    fo[1:] = np.cumsum([len(m) for m in mats])
    post = np.concatenate(mats) if fo[-1] else np.zeros((0, 5), mats[0].dtype)
    seqs, sc, cnt = oracle.beam_search_batch(post, fo, bw, tab, L, s_thr, r_thr, threads=THREADS)
    return ["".join("ACGT"[s] for s in q) for q in seqs], sc, cnt


@pytest.fixture(scope="module")
def table12():
    from radian_b200 import decode, synth

    tab = synth.make_table(12, 5)  # the table bench.py uses
    return tab, decode.RnaTable(tab)


def test_c3_headline_batch(table12):
    """bench.py's workload in small: 64 reads of the LogNormal length distribution plus one 10 kb
    read (430 k frames), bw 16, L 12, thresholds 0.5/0.5, float32."""
    from radian_b200 import decode, synth

    tab, lm = table12
    nb = np.concatenate([synth.read_lengths(64, 3000), [10000]])
    post, off = synth.make_reads(nb, seed=3 * 7919)
    mats = split(post.numpy(), off.numpy())
    seqs, scores, cnt = decode.beam_search_batch(mats, 16, lm, 0.5, 0.5, 12, return_details=True)
    want, wsc, wcnt = oracle_batch(mats, 16, tab, 12)
    assert seqs == want
    for i in range(len(mats)):
        assert close(scores[i, 0], wsc[i]), i
    assert np.array_equal(cnt[:, :2], wcnt)
    assert len(seqs[-1]) > 9000


def test_c4_chunk_1024_128_bw16():
    """config 4 as the reference runs chunk mode (basecall.py:110-123): every 1024-frame window at
    stride 128 decoded with the model off at bw 16, then stitched."""
    from oracle import oracle
    from radian_b200 import basecall, synth

    post, off = synth.make_reads(np.array([300, 1500, 40, 24]), seed=41)
    reads = split(post.numpy(), off.numpy())
    chunk_lists = [synth.split_windows(r, 1024, 128) for r in reads]
    args = types.SimpleNamespace(decode_type="chunk", beam_width=16, step_size=128)
    got = basecall.basecall_batch(list("abcd"), chunk_lists, args, None)
    flat = [m for cl in chunk_lists for m in cl]
    frags, _, _ = oracle_batch(flat, 16, None, 0)
    k = 0
    for cl, g in zip(chunk_lists, got):
        assert g == oracle.stitch(frags[k:k + len(cl)])[0]
        k += len(cl)
    assert len(got[1]) > 1000


def test_c4_global_1024_128_bw16(table12):
    """config 4 with the overlap merge (basecall.py:99-109): assemble_matrices over the same windows,
    then the RNA-LM decode at bw 16, L 12.  The single-window read stays float32 (matrix_assembly.py
    dtype rule) and must be decoded on the float32 path, the others on the float64 path."""
    from oracle import oracle
    from radian_b200 import basecall, synth

    tab, lm = table12
    post, off = synth.make_reads(np.array([300, 1500, 40, 20]), seed=42)
    reads = split(post.numpy(), off.numpy())
    chunk_lists = [synth.split_windows(r, 1024, 128) for r in reads]
    assert len(chunk_lists[3]) == 1  # 20 bases: fewer than 1024 frames
    args = types.SimpleNamespace(decode_type="global", beam_width=16, step_size=128, sig_threshold=0.5,
                                 rna_threshold=0.5, context_len=12)
    got = basecall.basecall_batch(list("abcd"), chunk_lists, args, lm)
    for cl, g in zip(chunk_lists, got):
        mat = oracle.assemble(cl, 128)
        seq, _, _, _ = oracle.beam_search(mat, 16, tab, 12, 0.5, 0.5, topk=1)
        assert g == "".join("ACGT"[s] for s in seq)
    assert oracle.assemble(chunk_lists[3], 128).dtype == np.float32
    assert oracle.assemble(chunk_lists[1], 128).dtype == np.float64


@pytest.mark.parametrize("bw", [6, 64])
@pytest.mark.parametrize("L", [6, 12])
def test_c5_corners(bw, L, table12):
    """Sweep corners: one 0.5 kb and one 10 kb read per cell (float32, thresholds 0.5/0.5)."""
    from radian_b200 import decode, synth

    if L == 12:
        tab, lm = table12
    else:
        tab = synth.make_table(L, 5)
        lm = decode.RnaTable(tab)
    post, off = synth.make_reads(np.array([500, 10000]), seed=500 + bw + L)
    mats = split(post.numpy(), off.numpy())
    seqs, scores, cnt = decode.beam_search_batch(mats, bw, lm, 0.5, 0.5, L, return_details=True)
    want, wsc, wcnt = oracle_batch(mats, bw, tab, L)
    assert seqs == want
    assert close(scores[0, 0], wsc[0]) and close(scores[1, 0], wsc[1])
    assert np.array_equal(cnt[:, :2], wcnt)


def tiny_matrix(rng, T, lo_exp, spread_every):
    """float64 'posteriors' whose entries go down to 10**lo_exp: every frame is scaled by a random
    power of ten, and every `spread_every` frames three of the four bases are a further 1e-250 below
    the rest, which pushes most candidates hundreds of orders of magnitude below the best one."""
    p = rng.dirichlet(np.ones(5), size=T)
    p *= 10.0 ** rng.uniform(lo_exp, 0, size=(T, 1))
    for t in range(0, T, spread_every):
        keep = rng.integers(0, 4)
        for c in range(4):
            if c != keep:
                p[t, c] *= 1e-250
    return p


@pytest.mark.parametrize("bw", [6, 16, 64])
def test_float64_dynamic_range(bw):
    """The reference scores in the log domain and therefore accepts any positive float64; the
    kernel's linear-domain scores are rescaled per read and must either agree with the oracle or say
    RADIAN_READ_RANGE (-> FloatingPointError), never return something else silently."""
    from radian_b200 import decode, synth

    rng = np.random.default_rng(bw)
    tab = synth.make_table(4, 9)
    lm = decode.RnaTable(tab)
    mats = [tiny_matrix(rng, 120, -300, 7), tiny_matrix(rng, 300, -250, 5), tiny_matrix(rng, 90, -40, 3),
            tiny_matrix(rng, 200, -300, 2)]
    want, wsc, _ = oracle_batch(mats, bw, tab, 4)
    n_ok = 0
    for i, m in enumerate(mats):
        try:
            seqs, scores, _ = decode.beam_search_batch([m], bw, lm, 0.5, 0.5, 4, return_details=True)
        except FloatingPointError:
            continue
        assert seqs[0] == want[i], i
        assert close(scores[0, 0], wsc[i]), i
        n_ok += 1
    assert n_ok >= 2  # the moderate cases must be inside the supported range


def test_float64_range_is_wide():
    """Entries down to 1e-300 on every frame (best beam loses ~500 bits per frame) with ~2^-500
    between a kept candidate and the next class of candidates, i.e. beams 2^-1000 apart in the first
    frames: inside the supported range, exact."""
    from radian_b200 import decode

    rng = np.random.default_rng(7)
    T = 64
    p = np.full((T, 5), 1e-300)
    p[:, 4] = 1e-160
    p[np.arange(T), rng.integers(0, 4, T)] = 1e-150  # one base 2^-498 above the rest
    p[::3, 4] = 1e-151
    want, wsc, _ = oracle_batch([p], 16, None, 0)
    seqs, scores, _ = decode.beam_search_batch([p], 16, None, None, None, None, return_details=True)
    assert seqs[0] == want[0]
    assert close(scores[0, 0], wsc[0])


def tie_read(T, seed, dtype):
    """A read whose first base is an exact two-way tie (P_A == P_C bit for bit): the labelings 'A...'
    and 'C...' are multiplied by the same factors ever after and keep bit-equal scores for the rest of
    the read (oracle == reference on these reads was checked in the build container)."""
    rng = np.random.default_rng(seed)
    p = np.zeros((T, 5), dtype)
    p[:, 4] = 0.9
    p[:, :4] = 0.025
    p[5] = [0.3, 0.3, 0.0, 0.0, 0.4]
    t = 12
    while t < T - 3:
        c = rng.integers(0, 4)
        p[t] = 0.01
        p[t, c] = 0.95
        p[t, 4] = 0.02
        t += rng.integers(4, 12)
    return p


@pytest.mark.parametrize("bw", [2, 6, 16, 40, 100])
def test_persistent_exact_ties(bw):
    """Kept beams with bit-equal scores over thousands of frames: ordered by the reference's dict
    insertion positions (stable sort, decode.py:35-39) in every frame, including the frames that only
    update the scores."""
    from radian_b200 import decode

    mats = [tie_read(3000, s, np.float32 if s % 2 else np.float64) for s in range(6)]
    want, wsc, _ = oracle_batch([m for m in mats if m.dtype == np.float32], bw, None, 0)
    want64, wsc64, _ = oracle_batch([m for m in mats if m.dtype == np.float64], bw, None, 0)
    seqs, scores, _ = decode.beam_search_batch(mats, bw, None, None, None, None, return_details=True)
    assert [s for s, m in zip(seqs, mats) if m.dtype == np.float32] == want
    assert [s for s, m in zip(seqs, mats) if m.dtype == np.float64] == want64
    assert all(len(s) > 300 for s in seqs)


def test_resident_batch_with_one_long_read(table12):
    """The resident entry point on a batch that one long read dominates: the launch planner gives that
    read a warp to itself and fewer CTAs per SM (csrc/decode.cu, plan_launch); results as ever."""
    import torch

    from radian_b200 import decode, synth

    tab, lm = table12
    nb = np.concatenate([np.full(40, 200), [6000], np.full(7, 900)])
    post, off = synth.make_reads(nb, seed=91, device="cuda")
    T = off[1:] - off[:-1]
    order = torch.argsort(T, descending=True).to(torch.int32)
    res = decode.decode_batch_device(post, off, 16, lm, 0.5, 0.5, max_frames=int(T.max()), order=order, counters=True)
    torch.cuda.synchronize()
    assert int(res.status.abs().sum()) == 0
    got = res.strings()
    p, o = post.cpu().numpy(), off.cpu().numpy()
    want, wsc, wcnt = oracle_batch(split(p, o), 16, tab, 12)
    assert got == want
    assert np.array_equal(res.counters.cpu().numpy()[:, :2].astype(np.uint64), wcnt)
    sc = res.scores.cpu().numpy()
    assert all(close(sc[i, 0], wsc[i]) for i in range(len(nb)))
