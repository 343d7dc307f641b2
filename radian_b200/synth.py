"""Synthetic workloads for tests and bench.py.

Nothing here is on the product path.  Two generators:

* ``make_table(L, seed)``   dense ``(4**L, 4)`` float64 RNA k-mer table.  Built only from
  integer hashing (splitmix64) and IEEE +,*,/ so that the same seed gives the same bits
  on every host -- golden fixtures can therefore reference a table by (L, seed).
* ``make_reads(...)``       posterior matrices of the read-length distribution named in
  BASELINE.json (LogNormal(median 1300, sigma 0.6) bases clipped to [200, 10000], about
  43 frames per base), as torch tensors on any device.

The real ``rnamodel_12mer_pc`` table and the sig2seq posteriors are not available
(SURVEY.md F2), so every number produced from these generators is "synthetic".
"""
from __future__ import annotations

import math

import numpy as np

_MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    """Vectorised splitmix64 finaliser on uint64 arrays (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _MASK
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _MASK
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _MASK
        x = x ^ (x >> np.uint64(31))
    return x


def _u01(h: np.ndarray) -> np.ndarray:
    """53-bit uniform in (0, 1) from a uint64 hash; exact in float64."""
    return ((h >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def make_table(L: int, seed: int = 0, chunk: int = 1 << 22) -> np.ndarray:
    """Dense (4**L, 4) float64 table; row i is the distribution after context index i.

    Context index = big-endian base 4 of the context, oldest symbol most significant,
    A,C,G,T = 0..3 (reference: basecall.py:54-57).  Rows come in three sharpness classes
    (w = u**16, u**4, u) so that roughly a quarter of the contexts have entropy < 0.5 nats,
    i.e. pass the reference's default ``--rna-threshold`` (decode.py:93).
    """
    n = 4 ** L
    out = np.empty((n, 4), dtype=np.float64)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        idx = np.arange(lo, hi, dtype=np.uint64)
        base = _splitmix64(idx ^ np.uint64((int(seed) * 0xD1342543DE82EF95) & 0xFFFFFFFFFFFFFFFF))
        cls = (base % np.uint64(10)).astype(np.int64)  # 0-2 sharp, 3-6 medium, 7-9 flat
        w = np.empty((hi - lo, 4), dtype=np.float64)
        for j in range(4):
            w[:, j] = _u01(_splitmix64(base + np.uint64(j + 1)))
        sharp = cls < 3
        med = (cls >= 3) & (cls < 7)
        w2 = w * w
        w4 = w2 * w2
        w16 = (w4 * w4) * (w4 * w4)
        w = np.where(sharp[:, None], w16, np.where(med[:, None], w4, w))
        s = ((w[:, 0] + w[:, 1]) + w[:, 2]) + w[:, 3]
        out[lo:hi] = w / s[:, None]
    return out


def table_entropy(table: np.ndarray) -> np.ndarray:
    """Vectorised entropy (nats) of each row; for statistics only, not bit-exact."""
    t = np.where(table > 0, table, 1.0)
    return -(table * np.log(t)).sum(axis=1)


def read_lengths(n_reads: int, seed: int, median: float = 1300.0, sigma: float = 0.6,
                 lo: int = 200, hi: int = 10000, fixed: int | None = None) -> np.ndarray:
    """Number of bases per read (LogNormal clipped, or a fixed length)."""
    if fixed is not None:
        return np.full(n_reads, int(fixed), dtype=np.int64)
    rng = np.random.default_rng([int(seed), 0x52414449])
    nb = np.exp(rng.normal(math.log(median), sigma, size=n_reads))
    return np.clip(nb, lo, hi).astype(np.int64)


def make_reads(n_bases, seed: int, device="cpu", zero_frac: float = 1e-3,
               frames_per_base: float = 43.0, dtype=None, second_spike: float = 0.3):
    """Synthetic posteriors for a batch of reads, concatenated along time.

    Returns (post, frame_offsets): ``post`` is a (sum T, 5) float32 tensor of softmax rows
    with the blank in column 4 (reference: decode.py:124), ``frame_offsets`` an int64
    tensor of n_reads+1 row offsets.  Per base: dwell = max(2, floor(Gamma(4, fpb/4)))
    frames; N(0,1) logits with +6 on the blank; a 1-3 frame spike (+12,+10,+8) on the true
    base; with p = ``second_spike`` (0.3) a second spike (+11,+9,+7) on a random base: the knob for
    how ambiguous the posteriors are, i.e. for the fraction of frames whose base distribution has
    entropy above ``--sig-threshold`` and for how often the beam changes (bench.py
    reports the measured fractions); ``zero_frac`` of the base entries forced to exact 0 so that
    the -inf paths of decode.py:16-17 are used.
    """
    import torch

    dtype = dtype or torch.float32
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed) & 0x7FFFFFFFFFFFFFFF)
    n_bases = torch.as_tensor(np.asarray(n_bases), dtype=torch.int64, device=dev)
    n_reads = n_bases.numel()
    nb_tot = int(n_bases.sum().item())
    # dwell per base
    alpha = torch.full((nb_tot,), 4.0, device=dev, dtype=torch.float32)
    gam = torch._standard_gamma(alpha, generator=g) * (frames_per_base / 4.0)
    dwell = torch.clamp(torch.floor(gam), min=2).to(torch.int64)
    base_start = torch.cumsum(dwell, 0) - dwell
    t_tot = int(dwell.sum().item())
    read_of_base_end = torch.cumsum(n_bases, 0)
    frame_offsets = torch.zeros(n_reads + 1, dtype=torch.int64, device=dev)
    csum = torch.cumsum(dwell, 0)
    frame_offsets[1:] = csum[read_of_base_end - 1]

    logits = torch.randn((t_tot, 5), generator=g, device=dev, dtype=torch.float32)
    logits[:, 4] += 6.0

    def spikes(heights, prob, true_base):
        sel = torch.rand(nb_tot, generator=g, device=dev) < prob
        u = torch.rand(nb_tot, generator=g, device=dev)
        slen = 1 + (u > 0.6).to(torch.int64) + (u > 0.85).to(torch.int64)
        slen = torch.minimum(slen, dwell)
        off = torch.floor(torch.rand(nb_tot, generator=g, device=dev) *
                          (dwell - slen + 1).to(torch.float32)).to(torch.int64)
        off = torch.minimum(off, dwell - slen)
        for k, h in enumerate(heights):
            m = sel & (slen > k)
            rows = (base_start + off + k)[m]
            cols = true_base[m]
            logits.index_put_((rows, cols), torch.full_like(rows, h, dtype=torch.float32),
                              accumulate=True)

    true_base = torch.randint(0, 4, (nb_tot,), generator=g, device=dev)
    spikes((12.0, 10.0, 8.0), 2.0, true_base)
    other = torch.randint(0, 4, (nb_tot,), generator=g, device=dev)
    spikes((11.0, 9.0, 7.0), float(second_spike), other)

    post = torch.softmax(logits, dim=1)
    del logits
    if zero_frac > 0:
        z = torch.rand((t_tot, 4), generator=g, device=dev) < zero_frac
        post[:, :4].masked_fill_(z, 0.0)
    return post.to(dtype), frame_offsets


def split_windows(post_np: np.ndarray, window: int, step: int):
    """Cut a (T,5) matrix into the overlapping window matrices the reference's sig model
    would emit (preprocess.py:4-22 windows + basecall.py:96 trim): every window has
    ``window`` rows except the last, which is trimmed to the frames that exist."""
    T = post_np.shape[0]
    mats = []
    start = 0
    while start + window <= T:
        mats.append(post_np[start:start + window])
        start += step
    mats.append(post_np[start:])  # len in [0, window); pad >= 1 in the reference
    return mats
