"""Chunk-mode fragment stitching on the host.

Behaviour of the reference's ``simple_assembly`` / ``index2base`` (radian/sequence_assembly.py:19-48,
90-97; call site basecall.py:122-123): consecutive fragments are aligned on their longest common
block (difflib) and every consensus column takes a vote.  Host string work, kept on the CPU.
"""
from __future__ import annotations

import difflib

import numpy as np

_BASE = {"A": 0, "C": 1, "G": 2, "T": 3, "a": 0, "c": 1, "g": 2, "t": 3}


def _add_votes(votes: np.ndarray, start: int, fragment: str) -> None:
    if start < 0:
        fragment = fragment[-start:]
        start = 0
    for i, b in enumerate(fragment):
        votes[_BASE[b], start + i] += 1


def simple_assembly(fragments):
    """-> (4, length) vote counts, same values as the reference for the same fragments."""
    width = 1000
    votes = np.zeros([4, width])
    pos = 0
    length = 0
    for k, frag in enumerate(fragments):
        if k == 0:
            _add_votes(votes, 0, frag)
            continue
        sm = difflib.SequenceMatcher(None, fragments[k - 1], frag)
        block = max(sm.get_matching_blocks(), key=lambda x: x[2])
        shift = block[0] - block[1]
        if shift + pos + len(frag) > width:
            votes = np.pad(votes, ((0, 0), (0, 1000)), mode="constant", constant_values=0)
            width += 1000
        _add_votes(votes, pos + shift, frag)
        pos += shift
        length = max(length, pos + len(frag))
    return votes[:, :length]


def index2base(indices) -> str:
    return "".join("ACGT"[int(x)] for x in indices)
