"""Host-side model of the ranking shortcuts the decode kernels take (csrc/decode.cu: lane-parallel
counting on high words with the rank-sum proof; csrc/decode_wide.cu: bisection in the copies'
descending order + histogram + prefix sums).  The claim the kernels rest on: whenever the shortcut
accepts its result, the ranks are those of the reference's stable sort (score desc, dict insertion
position asc; decode.py:35-39, 145); whenever two candidates that the shortcut cannot order share a
high word, it refuses and the exact pass runs.  Pure numpy: runs without a GPU."""
import zlib

import numpy as np
import pytest


def exact_ranks(keys, pos):
    """rank = number of candidates that sort before: (64-bit score desc, insertion position asc)."""
    n = len(keys)
    r = np.zeros(n, np.int64)
    for i in range(n):
        r[i] = np.sum((keys > keys[i]) | ((keys == keys[i]) & (pos < pos[i])))
    return r


def wide_shortcut(copy_keys, copy_rank, ext_keys):
    """decode_wide.cu, incremental ranks.  copy_rank: the copies' exact order among themselves (known
    from the order check on all 64 bits).  Returns (ok, copy_final_rank, ext_final_rank)."""
    na, ne = len(copy_keys), len(ext_keys)
    ch = copy_keys >> np.uint64(32)
    eh = ext_keys >> np.uint64(32)
    skey = np.zeros(na, np.uint64)
    skey[copy_rank] = ch                      # the copies' high words in rank order: descending
    assert np.all(skey[:-1] >= skey[1:])
    hist = np.zeros(na + 1, np.int64)
    ext_rank = np.zeros(ne, np.int64)
    tie = False
    nb = 1
    while nb < max(na, 1):
        nb *= 2
    for e in range(ne):
        c, step = 0, nb
        while step > 0:                       # branch-free bisection: largest c with skey[c - 1] > k
            t = c + step
            if t <= na and skey[min(t, na) - 1] > eh[e]:
                c = t
            step >>= 1
        if c < na:
            tie = tie or skey[c] == eh[e]     # a copy with my high word
            hist[c] += 1
        ext_rank[e] = c + np.sum(eh > eh[e])
    pre = np.cumsum(hist)                     # extensions with at most r copies above them
    copy_final = copy_rank + pre[copy_rank]
    mv = na + ne
    ok = (not tie) and (copy_final.sum() + ext_rank.sum() == mv * (mv - 1) // 2)
    return ok, copy_final, ext_rank


def narrow_shortcut(copy_keys, copy_rank, ext_keys):
    """decode.cu, lane-parallel counting on high words + the rank-sum proof."""
    ch = copy_keys >> np.uint64(32)
    eh = ext_keys >> np.uint64(32)
    copy_final = copy_rank + np.array([np.sum(eh > h) for h in ch], np.int64)
    ext_rank = np.array([np.sum(ch > h) + np.sum(eh > h) for h in eh], np.int64)
    mv = len(ch) + len(eh)
    ok = copy_final.sum() + ext_rank.sum() == mv * (mv - 1) // 2
    return ok, copy_final, ext_rank


def make_case(rng, na, ne, tie_kind):
    """Random positive float64 scores as bit patterns; tie_kind plants equal high words."""
    vals = np.exp(rng.uniform(-40, 0, na + ne))
    keys = vals.view(np.uint64).copy()
    if tie_kind == "ext-ext" and ne >= 2:
        i, j = rng.choice(ne, 2, replace=False)
        keys[na + j] = (keys[na + i] & np.uint64(0xFFFFFFFF00000000)) | np.uint64(rng.integers(1, 1 << 31))
    if tie_kind == "ext-copy" and ne >= 1 and na >= 1:
        i, j = rng.integers(na), rng.integers(ne)
        keys[na + j] = (keys[i] & np.uint64(0xFFFFFFFF00000000)) | np.uint64(rng.integers(1, 1 << 31))
    if tie_kind == "copy-copy" and na >= 2:
        i, j = rng.choice(na, 2, replace=False)
        keys[j] = (keys[i] & np.uint64(0xFFFFFFFF00000000)) | np.uint64(rng.integers(1, 1 << 31))
    if tie_kind == "copy-copy-exact" and na >= 2:
        i, j = rng.choice(na, 2, replace=False)
        keys[j] = keys[i]
    # dict insertion positions: copy of rank r at 5 r, extension (r, c) at 5 r + 1 + c (decode.cu)
    pos = np.zeros(na + ne, np.int64)
    order = np.argsort(-keys[:na].astype(np.float64), kind="stable")
    prov = np.empty(na, np.int64)
    prov[order] = np.arange(na)               # previous ranks: any permutation does for the positions
    pos[:na] = 5 * prov
    slots = rng.choice(max(na, 1) * 4, ne, replace=False)
    pos[na:] = 5 * (slots // 4) + 1 + slots % 4
    return keys, pos


@pytest.mark.parametrize("shortcut", [wide_shortcut, narrow_shortcut])
@pytest.mark.parametrize("tie_kind", ["none", "ext-ext", "ext-copy", "copy-copy", "copy-copy-exact"])
def test_shortcut_is_exact_or_refuses(shortcut, tie_kind):
    rng = np.random.default_rng(zlib.crc32((shortcut.__name__ + tie_kind).encode()))
    accepted = 0
    for _ in range(300):
        na = int(rng.integers(1, 129))
        ne = int(rng.integers(0, min(4 * na, 64) + 1))
        keys, pos = make_case(rng, na, ne, tie_kind)
        want = exact_ranks(keys, pos)
        copy_rank = exact_ranks(keys[:na], pos[:na])  # what the 64-bit order check / re-rank establishes
        ok, cf, ef = shortcut(keys[:na], copy_rank, keys[na:])
        if ok:
            accepted += 1
            assert np.array_equal(cf, want[:na]) and np.array_equal(ef, want[na:])
        elif tie_kind in ("none", "copy-copy", "copy-copy-exact"):
            # nothing the shortcut cannot order: it must not refuse (performance claim, not correctness)
            hi = keys >> np.uint64(32)
            ext_hi = hi[na:]
            clash = len(np.unique(ext_hi)) < ne or np.intersect1d(ext_hi, hi[:na]).size > 0
            assert clash, "refused without any high-word clash that involves an extension"
    if tie_kind == "none":
        assert accepted >= 290
    if tie_kind in ("ext-ext", "ext-copy"):
        assert accepted <= 150  # (cases without extensions, or too few of them to plant the tie)
