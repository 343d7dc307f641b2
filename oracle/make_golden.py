"""TEST INFRASTRUCTURE -- generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden [--long]

Every fixture stores the exact inputs handed to the reference's own
``decode.beam_search`` / ``matrix_assembly.assemble_matrices`` and what they returned
(plus the final candidates' pr_total values captured by wrapping
``BeamList.sort_labelings``, see ref_loader.beam_search_with_scores).  Large RNA tables
are stored as (L, seed) of radian_b200.synth.make_table, which is bit-reproducible.
Environment of the recorded run: CPython 3.12, numpy 2.3.5, scikit-learn 1.9.0
(the reference pins numpy~=1.19.5 / scikit-learn~=1.1.2, requirements.txt:5,9).
"""
from __future__ import annotations

import argparse
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from radian_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
SYM = {"A": 0, "C": 1, "G": 2, "T": 3}


def seq_to_u8(s):
    return np.array([SYM[c] for c in s], dtype=np.uint8)


def random_posteriors(rng, T, peaky, zero_frac, dtype):
    """Small random softmax matrices with optional exact zeros (not renormalised)."""
    logits = rng.normal(0, 1, size=(T, 5))
    logits[:, 4] += rng.choice([0.0, 2.0, 6.0])
    if peaky:
        k = rng.integers(0, 4, size=T)
        m = rng.random(T) < 0.3
        logits[np.arange(T)[m], k[m]] += rng.choice([4.0, 8.0, 12.0])
    p = np.exp(logits - logits.max(1, keepdims=True))
    p = (p / p.sum(1, keepdims=True)).astype(np.float32)
    if zero_frac:
        z = rng.random((T, 5)) < zero_frac
        p[z] = 0.0
    return p.astype(dtype)


_TABLES = {}


def run_case(args):
    mat, bw, L, tseed, s_thr, r_thr = args
    dec, _, _ = ref_loader.load()
    lm = None
    if L:
        if (L, tseed) not in _TABLES:
            _TABLES.clear()
            _TABLES[(L, tseed)] = synth.make_table(L, tseed)
        lm = ref_loader.DenseLM(_TABLES[(L, tseed)])
    t0 = time.time()
    seq, scores, ncomb = ref_loader.beam_search_with_scores(dec, mat, bw, lm, s_thr, r_thr, L, topk=8)
    nl = lm.n_lookup if lm is not None else 0
    return seq, scores, nl, ncomb, time.time() - t0


def pack_cases(cases, results):
    """Flatten variable-length cases into arrays for one npz."""
    out = {}
    mats32 = [c[0] for c in cases if c[0].dtype == np.float32]
    mats64 = [c[0] for c in cases if c[0].dtype == np.float64]
    out["post32"] = np.concatenate(mats32) if mats32 else np.zeros((0, 5), np.float32)
    out["post64"] = np.concatenate(mats64) if mats64 else np.zeros((0, 5), np.float64)
    meta = []
    o32 = o64 = 0
    seqs = []
    scores = np.full((len(cases), 8), np.nan)
    for i, (c, r) in enumerate(zip(cases, results)):
        mat, bw, L, tseed, s_thr, r_thr = c
        T = mat.shape[0]
        is64 = int(mat.dtype == np.float64)
        off = o64 if is64 else o32
        if is64:
            o64 += T
        else:
            o32 += T
        seq, sc, nl, ncomb, _ = r
        seqs.append(seq_to_u8(seq))
        scores[i, :len(sc)] = sc
        meta.append([is64, off, T, bw, L, tseed, len(seq), nl, ncomb])
    out["meta"] = np.array(meta, dtype=np.int64)
    out["thr"] = np.array([[c[4] if c[4] is not None else np.nan,
                            c[5] if c[5] is not None else np.nan] for c in cases])
    out["seq"] = np.concatenate(seqs) if seqs else np.zeros(0, np.uint8)
    out["scores"] = scores
    return out


def gen_decode_random(n_cases, seed):
    rng = np.random.default_rng(seed)
    cases = []
    for i in range(n_cases):
        T = int(rng.integers(1, 400))
        bw = int(rng.choice([1, 2, 3, 6, 16, 32, 64]))
        lm_on = rng.random() < 0.7
        L = int(rng.integers(1, 7)) if lm_on else 0
        dtype = np.float64 if rng.random() < 0.6 else np.float32
        peaky = rng.random() < 0.7
        zf = float(rng.choice([0.0, 0.0, 0.01, 0.2]))
        mat = random_posteriors(rng, T, peaky, zf, dtype)
        if lm_on:
            s_thr = float(rng.choice([0.0, 0.5, 0.5, 1.0, np.log(4.0), 0.3]))
            r_thr = float(rng.choice([0.0, 0.5, 0.5, 1.0, np.log(4.0), 0.9]))
        else:
            s_thr = r_thr = None
        cases.append((mat, bw, L, int(rng.integers(0, 1 << 30)) if lm_on else 0, s_thr, r_thr))
    return cases


def gen_decode_kat():
    """Known-answer micro-cases from SURVEY.md section 4."""
    cases = []
    blank = np.zeros((7, 5), np.float32)
    blank[:, 4] = 1.0
    cases.append((blank, 6, 0, 0, None, None))                       # all blank -> ''
    two = np.zeros((6, 5), np.float32)
    two[:, 1] = 0.5
    two[:, 4] = 0.5
    cases.append((two, 4, 0, 0, None, None))                         # exact zeros, -inf beams kept
    cases.append((np.full((3, 5), 0.2, np.float32), 3, 0, 0, None, None))   # stable tie-break -> 'A'
    cases.append((np.full((9, 5), 0.2, np.float64), 6, 2, 7, 0.5, 0.5))     # ties with the LM on
    cases.append((np.full((5, 5), 0.2, np.float32), 16, 0, 0, None, None))
    z = np.zeros((4, 5), np.float32)                                  # all-zero rows
    cases.append((z, 3, 0, 0, None, None))
    one = np.zeros((1, 5), np.float32)
    one[0] = [0.7, 0.1, 0.1, 0.05, 0.05]
    cases.append((one, 1, 0, 0, None, None))                          # T=1, bw=1
    rep = np.zeros((8, 5), np.float64)                                # repeats need a blank between
    rep[:, 0] = [0.9, 0.05, 0.9, 0.9, 0.05, 0.9, 0.05, 0.05]
    rep[:, 4] = 1 - rep[:, 0]
    cases.append((rep, 6, 1, 3, 0.0, 2.0))
    return cases


def gen_decode_real_shape(quick):
    """Reads drawn from the bench generator (synth.make_reads), default flags of the reference."""
    import torch  # noqa: F401

    cases = []
    nb = np.array([60, 35, 90, 120] if quick else [60, 35, 90, 120, 150, 100])
    post, off = synth.make_reads(nb, seed=11)
    post = post.numpy()
    off = off.numpy()
    for i in range(len(nb)):
        m32 = post[off[i]:off[i + 1]]
        # config 2: pure CTC, bw=6, float32 (chunk-mode call, basecall.py:113-120)
        cases.append((m32.copy(), 6, 0, 0, None, None))
        # config 3: LM global decode, bw=16, L in {6, 11}, thresholds 0.5/0.5, float64 matrix
        m64 = m32.astype(np.float64)
        m64 = m64 / np.abs(m64).sum(1, keepdims=True)
        cases.append((m64, 16, 6 if i % 2 else 11, 5, 0.5, 0.5))
    # config 1 shape: bw=6 (default), L=11, float64
    cases.append((post[off[0]:off[1]].astype(np.float64), 6, 11, 5, 0.5, 0.5))
    return cases


def gen_assembly(n_cases, seed):
    _, ma, _ = ref_loader.load()
    rng = np.random.default_rng(seed)
    recs = []
    chunks_all, lens_all, outs32, outs64 = [], [], [], []
    for i in range(n_cases):
        W = int(rng.integers(1, 80))
        S = int(rng.integers(1, W + 1))
        if i % 4 == 0:      # the reference's own windowing (preprocess.py:4-22 + basecall.py:96)
            T = int(rng.integers(1, 6 * W))
            full = random_posteriors(rng, T, True, 0.02, np.float32)
            if i % 8 == 0:
                full[rng.integers(0, T)] = 0.0    # an all-zero row: norm 0 -> left unchanged
            mats = synth.split_windows(full, W, S)
        else:               # ragged chunk lengths (anything the reference accepts)
            n = int(rng.integers(1, 7))
            mats = []
            end = 0
            for k in range(n):
                lo = max(0, k * S - end)  # must at least reach the current end (no gaps)
                ln = int(rng.integers(lo, lo + W + 1)) if k * S <= end else 0
                if k * S > end:
                    ln = 0
                mats.append(random_posteriors(rng, ln, True, 0.05, np.float32) if ln else
                            np.zeros((0, 5), np.float32))
                if ln:
                    end = max(end, k * S + ln)
            if not any(len(m) for m in mats):
                mats[0] = random_posteriors(rng, 3, True, 0.0, np.float32)
        ref = ma.assemble_matrices([m for m in mats], S)
        ref = np.asarray(ref)
        if ref.ndim == 1:
            ref = ref.reshape(0, 5)
        lens = [len(m) for m in mats]
        recs.append([S, len(mats), sum(lens), ref.shape[0], int(ref.dtype == np.float64)])
        lens_all.extend(lens)
        chunks_all.extend([m for m in mats if len(m)])
        (outs64 if ref.dtype == np.float64 else outs32).append(ref)
    return {
        "meta": np.array(recs, dtype=np.int64),
        "chunk_lens": np.array(lens_all, dtype=np.int32),
        "chunks": np.concatenate(chunks_all),
        "out32": np.concatenate(outs32).astype(np.float32) if outs32 else np.zeros((0, 5), np.float32),
        "out64": np.concatenate(outs64) if outs64 else np.zeros((0, 5), np.float64),
    }


def gen_sequence_assembly(n_cases, seed):
    """Chunk-mode stitching (sequence_assembly.py:19-48,90-97) on overlapping noisy fragments."""
    _, _, sa = ref_loader.load()
    rng = np.random.default_rng(seed)
    cases = []
    for _ in range(n_cases):
        n = int(rng.integers(40, 400))
        truth = "".join("ACGT"[i] for i in rng.integers(0, 4, n))
        frags = []
        start = 0
        while start < n:
            ln = int(rng.integers(15, 40))
            f = list(truth[start:start + ln])
            for j in range(len(f)):
                if rng.random() < 0.03:
                    f[j] = "ACGT"[rng.integers(0, 4)]
            frags.append("".join(f))
            start += int(rng.integers(3, 12))
        if len(frags) == 1 or rng.random() < 0.1:
            frags = frags[:1]
        cons = sa.simple_assembly(frags)
        cases.append({"fragments": frags, "consensus": sa.index2base(np.argmax(cons, axis=0)),
                      "votes_shape": list(cons.shape), "votes_sum": float(cons.sum())})
    return cases


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--long", action="store_true", help="also the slow full-length reads")
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    a = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)

    np.savez_compressed(os.path.join(GOLDEN, "assembly.npz"), **gen_assembly(120, 2024))
    print("assembly.npz written")
    import json

    with open(os.path.join(GOLDEN, "sequence_assembly.json"), "w") as f:
        json.dump(gen_sequence_assembly(40, 5), f)
    print("sequence_assembly.json written")

    sets = {
        "decode_kat": gen_decode_kat(),
        "decode_random": gen_decode_random(240, 77),
        "decode_synth": gen_decode_real_shape(quick=not a.long),
    }
    if a.long:
        post, off = synth.make_reads(np.array([1500]), seed=12)
        m32 = post.numpy()
        m64 = m32.astype(np.float64)
        m64 = m64 / np.abs(m64).sum(1, keepdims=True)
        sets["decode_long"] = [(m32, 6, 0, 0, None, None), (m64, 16, 11, 5, 0.5, 0.5)]
    with Pool(a.procs) as pool:
        for name, cases in sets.items():
            t0 = time.time()
            res = pool.map(run_case, cases, chunksize=1)
            np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"), **pack_cases(cases, res))
            print(f"{name}.npz: {len(cases)} cases, {time.time() - t0:.1f}s wall, "
                  f"ref cpu {sum(r[4] for r in res):.1f}s")


if __name__ == "__main__":
    main()


def gen_sequence_assembly_ext(seed=9):
    """More stitching cases for the GPU path: long fragments (difflib's
    'popular element' rule at len(b) >= 200), unrelated and empty fragments, drifting
    positions, and the reference's IndexError when a fragment outgrows the vote buffer.
    Run on its own:  python -c "from oracle import make_golden as m; m.write_sequence_assembly_ext()" """
    _, _, sa = ref_loader.load()
    rng = np.random.default_rng(seed)

    def noisy(s, p):
        f = list(s)
        for j in range(len(f)):
            if rng.random() < p:
                f[j] = "ACGT"[rng.integers(0, 4)]
        return "".join(f)

    def rand_seq(n, probs=None):
        return "".join("ACGT"[i] for i in rng.choice(4, n, p=probs))

    lists = []
    for _ in range(60):  # typical chunk-mode fragments
        n = int(rng.integers(30, 500))
        truth = rand_seq(n)
        frags, start = [], 0
        while start < n:
            frags.append(noisy(truth[start:start + int(rng.integers(10, 45))], 0.04))
            start += int(rng.integers(1, 14))
        lists.append(frags)
    for _ in range(12):  # long fragments: every base is 'popular' in b
        n = int(rng.integers(600, 1500))
        truth = rand_seq(n)
        frags, start = [], 0
        while start < n:
            frags.append(noisy(truth[start:start + int(rng.integers(180, 330))], 0.02))
            start += int(rng.integers(20, 120))
        lists.append(frags)
    for _ in range(8):  # long fragments with one or two rare letters (mixed popular / not)
        n = int(rng.integers(500, 900))
        truth = list(rand_seq(n, [0.5, 0.5, 0.0, 0.0]))
        for j in rng.choice(n, size=max(2, n // 90), replace=False):
            truth[j] = "GT"[rng.integers(0, 2)]
        truth = "".join(truth)
        frags, start = [], 0
        while start < n:
            frags.append(truth[start:start + int(rng.integers(190, 260))])
            start += int(rng.integers(30, 90))
        lists.append(frags)
    for _ in range(10):  # unrelated / empty / tiny fragments
        k = int(rng.integers(1, 9))
        frags = []
        for _ in range(k):
            r = rng.random()
            frags.append("" if r < 0.2 else rand_seq(int(rng.integers(1, 4))) if r < 0.4 else rand_seq(int(rng.integers(5, 40))))
        lists.append(frags)
    lists.append(["ACGT"])
    lists.append([])
    lists.append(["", ""])
    lists.append(["AAAAAAAAAA", "AAAAAAAAAA", "AAAAA", "CAAAAAAAAAAG"])
    lists.append([rand_seq(30)] + [rand_seq(12) + "GATTACAGATTACA" + rand_seq(3) for _ in range(5)])  # drifts left
    lists.append([rand_seq(1200), rand_seq(20)])      # first fragment outgrows the 1000-column buffer
    lists.append([rand_seq(400), rand_seq(700), rand_seq(900)])
    big = rand_seq(2600)
    lists.append([big[i:i + 300] for i in range(0, 2300, 100)])  # buffer growth several times
    cases = []
    for frags in lists:
        try:
            cons = sa.simple_assembly(frags)
            cases.append({"fragments": frags, "consensus": sa.index2base(np.argmax(cons, axis=0)) if cons.shape[1] else "",
                          "votes": cons.astype(int).tolist()})
        except Exception as e:  # the reference's own failure modes are part of its behaviour
            cases.append({"fragments": frags, "error": type(e).__name__})
    return cases


def write_sequence_assembly_ext():
    import json

    os.makedirs(GOLDEN, exist_ok=True)
    cases = gen_sequence_assembly_ext()
    with open(os.path.join(GOLDEN, "sequence_assembly_ext.json"), "w") as f:
        json.dump(cases, f)
    print("sequence_assembly_ext.json:", len(cases), "cases,", sum("error" in c for c in cases), "raising")


def write_preprocess_golden(seed=17):
    """Fixtures for SURVEY.md 8f N2 from the unmodified reference (preprocess.py:4-49): the five raw
    signals of the bundled radian/data/reads.fast5 (read with radian_b200/fast5.py) and synthetic
    int16 signals incl. ties, even/odd lengths, the int64 quirk of np.vectorize, and both
    ValueErrors.  Run on its own:
        python -c "from oracle import make_golden as m; m.write_preprocess_golden()" """
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_preprocess", "/root/reference/radian/preprocess.py")
    P = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(P)
    from radian_b200 import fast5

    rng = np.random.default_rng(seed)
    cases = []  # (signal int16, outlier (python int or float))
    for rid, sig in fast5.reads("/root/reference/radian/data/reads.fast5"):
        cases.append((sig, 4))
    for n in (1, 2, 3, 4, 5, 8, 31, 32, 33, 100, 255, 256, 257, 1000, 4097, 20000):
        base = rng.normal(700, 60, n)
        base[rng.random(n) < 0.02] += rng.normal(0, 900, int((rng.random(n) < 0.02).sum()) or 1)[0]
        cases.append((np.clip(np.rint(base), -32768, 32767).astype(np.int16), 4))
    for n in (10, 500, 3000):
        cases.append((rng.integers(-32768, 32768, n).astype(np.int16), 3))           # full range
        cases.append((rng.integers(690, 712, n).astype(np.int16), 2.5))              # many ties, float clip
        s = rng.integers(600, 800, n).astype(np.int16)
        s[0] = 30000                                                                 # first sample clipped: int64 result
        cases.append((s, 4))
        s = s.copy()
        s[0] = -30000
        cases.append((s, 4.0))                                                       # float clip: stays float64
    cases.append((np.full(50, 123, np.int16), 4))                                    # MAD zero
    cases.append((np.array([5, 5, 5, 6], np.int16), 4))                              # MAD zero (median of distances)
    cases.append((np.array([1, 2], np.int16), 4))
    cases.append((np.zeros(0, np.int16), 4))                                         # empty
    sig_all, off, outl, is_int, res, res_int, err = [], [0], [], [], [], [], []
    win = []
    for sig, o in cases:
        sig_all.append(sig)
        off.append(off[-1] + len(sig))
        outl.append(float(o))
        is_int.append(isinstance(o, int))
        try:
            r = P.mad_normalise(sig, o)
            err.append(0)
            res_int.append(r.dtype == np.int64)
            res.append(r.astype(np.int64).view(np.int64) if r.dtype == np.int64 else r.view(np.int64))
            for (W, S) in ((1024, 128), (64, 64), (50, 7)):
                w, pad = P.get_windows(r, W, S)
                win.append((len(err) - 1, W, S, w.shape[0], pad, float(w.sum()), float((w * np.arange(1, w.size + 1).reshape(w.shape) % 7).sum())))
        except ValueError as e:
            err.append(1 if "empty" in str(e) else 2)
            res_int.append(False)
            res.append(np.zeros(len(sig), np.int64))
    np.savez_compressed(os.path.join(GOLDEN, "preprocess.npz"), signal=np.concatenate(sig_all), offsets=np.array(off),
                        outlier=np.array(outl), outlier_is_int=np.array(is_int), result_bits=np.concatenate(res),
                        result_is_int=np.array(res_int), error=np.array(err), windows=np.array(win))
    print("preprocess.npz:", len(cases), "cases,", sum(e != 0 for e in err), "raising,", sum(res_int), "int64 results,",
          len(win), "window checks")


def write_decode_wide(seed=404):
    """Reference fixtures for the wide kernel (beam widths above 32): bench-style posteriors of
    short reads, model on and off, float32 and float64.  Run on its own:
        python -c "from oracle import make_golden as m; m.write_decode_wide()" """
    rng = np.random.default_rng(seed)
    cases = []
    for bw in (33, 48, 64, 65, 100, 128):
        for L in (0, 3, 5):
            nb = int(rng.integers(2, 6))
            post, _ = synth.make_reads(np.array([nb]), seed=int(rng.integers(1 << 30)))
            mat = post.numpy()
            if rng.random() < 0.5:
                mat = mat.astype(np.float64)
            cases.append((mat, bw, L, int(rng.integers(0, 1 << 30)) if L else 0, 0.5 if L else None,
                          0.5 if L else None))
    with Pool(os.cpu_count()) as pool:
        res = pool.map(run_case, cases, chunksize=1)
    np.savez_compressed(os.path.join(GOLDEN, "decode_wide.npz"), **pack_cases(cases, res))
    print("decode_wide.npz:", len(cases), "cases, ref cpu %.1fs" % sum(r[4] for r in res))


def write_decode_headline(seed=1212):
    """Reference fixtures on BASELINE.json's own configurations (they all use the 12-symbol
    context of the headline metric, which the older fixtures stop short of):
      C3  RNA-LM global decode, bw 16, L 12, thresholds 0.5/0.5, float32 posteriors as the bench
          feeds them, plus two float64 (assembled-matrix style) reads;
      C4  chunk-len 1024 / step-size 128, bw 16: (i) every window matrix decoded with the model off
          (basecall.py:110-120) and (ii) the reference's own assemble_matrices output decoded with
          the model on (basecall.py:99-109);
      C5  corners of the sweep that the reference can do in minutes: bw 6 and 64, L 6 and 12.
    Run on its own (8 processes, a few CPU-minutes; the L=12 table is 512 MB per process):
        python -c "from oracle import make_golden as m; m.write_decode_headline()" """
    _, ma, _ = ref_loader.load()
    rng = np.random.default_rng(seed)
    cases = []
    # C3
    nb = np.array([30, 55, 80, 100])
    post, off = synth.make_reads(nb, seed=int(rng.integers(1 << 30)))
    post, off = post.numpy(), off.numpy()
    for i in range(len(nb)):
        cases.append((post[off[i]:off[i + 1]].copy(), 16, 12, 5, 0.5, 0.5))
    for i in (0, 2):
        m64 = post[off[i]:off[i + 1]].astype(np.float64)
        cases.append((m64 / np.abs(m64).sum(1, keepdims=True), 16, 12, 5, 0.5, 0.5))
    # C4: one read of ~70 bases -> windows of 1024 frames every 128
    p4, _ = synth.make_reads(np.array([70]), seed=int(rng.integers(1 << 30)))
    mats = synth.split_windows(p4.numpy(), 1024, 128)
    for m in mats:
        if len(m):
            cases.append((np.ascontiguousarray(m), 16, 0, 0, None, None))
    glob = np.asarray(ma.assemble_matrices([m for m in mats], 128))
    assert glob.dtype == np.float64
    cases.append((glob, 16, 12, 5, 0.5, 0.5))
    # C5 corners
    p5, o5 = synth.make_reads(np.array([500, 500, 100, 100]), seed=int(rng.integers(1 << 30)))
    p5, o5 = p5.numpy(), o5.numpy()
    cases.append((p5[o5[0]:o5[1]].copy(), 6, 6, 5, 0.5, 0.5))
    cases.append((p5[o5[1]:o5[2]].copy(), 6, 12, 5, 0.5, 0.5))
    cases.append((p5[o5[2]:o5[3]].copy(), 64, 6, 5, 0.5, 0.5))
    cases.append((p5[o5[3]:o5[4]].copy(), 64, 12, 5, 0.5, 0.5))
    t0 = time.time()
    # longest first so that the pool stays balanced
    order = sorted(range(len(cases)), key=lambda i: -cases[i][0].shape[0] * cases[i][1])
    with Pool(min(8, os.cpu_count())) as pool:
        res_o = pool.map(run_case, [cases[i] for i in order], chunksize=1)
    res = [None] * len(cases)
    for i, r in zip(order, res_o):
        res[i] = r
    np.savez_compressed(os.path.join(GOLDEN, "decode_headline.npz"), **pack_cases(cases, res))
    print("decode_headline.npz:", len(cases), "cases, %.1fs wall, ref cpu %.1fs" % (time.time() - t0, sum(r[4] for r in res)))


def write_decode_sparse(seed=77):
    """Reference behaviour with a model that lacks contexts (a plain dict without some keys): the
    search raises KeyError at lm[context] (decode.py:83) as soon as a kept labeling reaches a missing
    context, and is unaffected otherwise.  -> tests/golden/decode_sparse.npz
        python -c "from oracle import make_golden as m; m.write_decode_sparse()" """
    dec, _, _ = ref_loader.load()
    rng = np.random.default_rng(seed)
    mats, meta, seqs, masks = [], [], [], []
    for i in range(40):
        T = int(rng.integers(3, 120))
        L = int(rng.integers(1, 4))
        bw = int(rng.choice([1, 3, 6, 16, 40]))
        tseed = int(rng.integers(1 << 20))
        frac = float(rng.choice([0.02, 0.1, 0.3, 0.9]))
        mat = random_posteriors(rng, T, True, 0.0, np.float32 if i % 2 else np.float64)
        if i % 5 == 0:  # strongly peaked on two symbols: few contexts are ever reached
            mat[:, 2:4] *= 1e-6
        tab = synth.make_table(L, tseed)
        present = rng.random(4 ** L) >= frac
        lm = {}
        for idx in np.flatnonzero(present):
            lm[tuple((int(idx) >> (2 * (L - 1 - j))) & 3 for j in range(L))] = [float(x) for x in tab[idx]]
        try:
            seq = dec.beam_search(mat, "ACGT", bw, lm, 0.5, 0.5, L, {})
            err = 0
        except KeyError:
            seq, err = "", 1
        mats.append(mat.astype(np.float64))
        meta.append([T, L, bw, tseed, err, len(seq), int(mat.dtype == np.float64)])
        seqs.append(seq_to_u8(seq))
        masks.append(present)
    np.savez_compressed(os.path.join(GOLDEN, "decode_sparse.npz"), post=np.concatenate(mats), meta=np.array(meta, dtype=np.int64),
                        seq=np.concatenate(seqs), present=np.concatenate(masks))
    print("decode_sparse.npz:", len(meta), "cases,", sum(m[4] for m in meta), "raising KeyError")
