"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* comprna/radian reference.

Only usable in the build container (``/root/reference`` does not exist on the GPU
box).  It is used by ``oracle/make_golden.py`` to produce the committed fixtures under
``tests/golden/`` and by the container-only cross-checks in ``tests/`` (skipped when the
reference tree is absent).  Nothing in the product package imports this file.

The reference's hot-path modules import two packages they never call on this path:
``tensorflow`` (annotation only, radian/decode.py:11,104) and ``matplotlib.pyplot``
(plot helpers, radian/matrix_assembly.py:1,55-77).  Both are replaced by empty stub
modules so that ``decode.py`` / ``matrix_assembly.py`` import byte-for-byte unmodified.
"""
import os
import sys
import types

REFERENCE_DIR = os.environ.get("RADIAN_REFERENCE_DIR", "/root/reference/radian")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "decode.py"))


def load():
    """Return (decode, matrix_assembly, sequence_assembly) modules of the reference."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_DIR}")
    if "tensorflow" not in sys.modules:
        tf = types.ModuleType("tensorflow")
        tf.keras = types.SimpleNamespace(Model=object)
        sys.modules["tensorflow"] = tf
    if "matplotlib.pyplot" not in sys.modules:
        mpl = sys.modules.get("matplotlib") or types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules.setdefault("matplotlib", mpl)
        sys.modules["matplotlib.pyplot"] = plt
    # import under private names so they never shadow the product's modules
    import importlib.util

    mods = []
    for name in ("decode", "matrix_assembly", "sequence_assembly"):
        key = f"_radian_reference_{name}"
        if key in sys.modules:
            mods.append(sys.modules[key])
            continue
        spec = importlib.util.spec_from_file_location(key, os.path.join(REFERENCE_DIR, f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[key] = mod
        spec.loader.exec_module(mod)
        mods.append(mod)
    return tuple(mods)


class DenseLM:
    """dict-like view of a dense (4**L, 4) float64 table for the unmodified reference.

    The reference only ever does ``if lm`` (decode.py:157,180) and ``model[context]``
    with a length-L tuple (decode.py:83), so this behaves exactly like the complete
    ``{tuple: [pA,pC,pG,pT]}`` dict built at basecall.py:50-57 without materialising
    4**L Python tuples.  Index = big-endian base 4, oldest symbol most significant.
    """

    def __init__(self, table):
        self.table = table
        self.n_lookup = 0

    def __bool__(self):
        return True

    def __getitem__(self, ctx):
        idx = 0
        for c in ctx:
            idx = idx * 4 + int(c)
        self.n_lookup += 1
        return [float(x) for x in self.table[idx]]


def beam_search_with_scores(decode, mat, beam_width, lm, s_thr, r_thr, L, topk=8):
    """Run the reference beam_search and also capture the final candidates' scores.

    ``beam_search`` returns only the string (decode.py:207-212).  The last BeamList whose
    ``sort_labelings`` is called is the final one (decode.py:207), so wrapping that method
    exposes the final (pr_total, labeling) list without editing the reference.
    Returns (sequence, [pr_total of the stable-sorted final candidates][:topk], n_combine).
    """
    seen = {}
    orig_sort = decode.BeamList.sort_labelings
    orig_comb = decode.combine_dists
    counter = {"combine": 0}

    def sort_spy(self):
        seen["last"] = self
        return orig_sort(self)

    def comb_spy(r, s):
        counter["combine"] += 1
        return orig_comb(r, s)

    decode.BeamList.sort_labelings = sort_spy
    decode.combine_dists = comb_spy
    try:
        seq = decode.beam_search(mat, "ACGT", beam_width, lm, s_thr, r_thr, L, {} if lm else None)
    finally:
        decode.BeamList.sort_labelings = orig_sort
        decode.combine_dists = orig_comb
    final = seen["last"]
    beams = sorted(final.entries.values(), reverse=True, key=lambda x: x.pr_total)
    scores = [float(b.pr_total) for b in beams][:topk]
    return seq, scores, counter["combine"]
