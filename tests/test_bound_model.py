"""The integer bound of the quiet-frame test (csrc/decode_common.cuh: kSlackPlain, kSlackGated;
csrc/decode.cu, "QUIET frames"), checked numerically on the host.

With h(x) = high word of the float64 x, (h(x) >> 20) - 1023 + mantissa fraction is a lower bound of
log2 x that is short by at most 0.0861, so
    h(p) + h(d) - bias + slack < h(w)                implies  p * d < w            (plain extension)
    h(p) + max(h(r), h(q)) + h(S) - 2 bias + slack' < h(w)
                                                     implies  p * ((r + q) * (S / 2)) < w   (gated)
A frame is only declared quiet on the strength of these implications; a counter-example would be a
wrong basecall.  Random and adversarial operands (mantissas where the linear bound of log2 is
worst, thresholds one unit above and below the bound)."""
import numpy as np

BIAS = 0x3FF00000
SLACK_PLAIN = 181000 - BIAS        # decode_common.cuh: kSlackPlain
SLACK_GATED = 272000 - BIAS        # decode_common.cuh: kSlackGated (one more factor, one more bias below)


def hi(x):
    return (np.asarray(x, np.float64).view(np.uint64) >> np.uint64(32)).astype(np.int64)


def from_hi(h, lo=0):
    return ((np.asarray(h, np.int64).astype(np.uint64) << np.uint64(32)) | np.uint64(lo)).view(np.float64)


def operands(rng, n, lo_exp, hi_exp):
    """positive normal doubles: random exponents, mantissas random or near 1 + (1/ln 2 - 1), where
    log2(1 + f) - f is largest"""
    e = rng.integers(lo_exp, hi_exp, n)
    f = np.where(rng.random(n) < 0.5, rng.random(n), 0.4427 + rng.normal(0, 0.01, n).clip(-0.05, 0.05))
    return np.ldexp(1.0 + f, e)


def test_plain_bound_never_lies():
    rng = np.random.default_rng(5)
    n = 2_000_000
    p = operands(rng, n, -900, 900)            # a beam's score, anywhere in the rescaled range
    d = operands(rng, n, -149, 0).clip(max=1)  # a probability
    ub = hi(p) + hi(d) + SLACK_PLAIN
    prod = p * d
    # worst copies right at the edge of the test: the smallest high word the test accepts, and the
    # largest one it refuses, with an all-zero low word (the hardest w for the implication)
    w_accept = from_hi(ub + 1)
    ok = (ub + 1 > 0x00100000) & (ub + 1 < 0x7FE00000)
    assert np.all(prod[ok] < w_accept[ok])
    # the slack is not wasteful either: dropping a quarter of it does produce counter-examples
    w_tight = from_hi(ub + 1 - 181000 // 4)
    assert np.any(prod[ok] >= w_tight[ok])


def test_gated_bound_never_lies():
    rng = np.random.default_rng(6)
    n = 2_000_000
    p = operands(rng, n, -900, 900)
    r = np.where(rng.random(n) < 0.1, 0.0, operands(rng, n, -40, 3).clip(max=4.0))   # table entries in [0, 4]
    q = operands(rng, n, -60, 0).clip(max=1)                                         # p_c / S
    S = operands(rng, n, -100, 0).clip(max=1)                                        # sum of the four bases
    emission = (r + q) * (S * 0.5)             # combine_dists as the kernels compute it
    ub = hi(p) + np.maximum(hi(r), hi(q)) + hi(S) - BIAS + SLACK_GATED
    prod = p * emission
    w_accept = from_hi(ub + 1)
    ok = (ub + 1 > 0x00100000) & (ub + 1 < 0x7FE00000)
    assert ok.sum() > n // 2
    assert np.all(prod[ok] < w_accept[ok])
