"""Command line driver with the reference's flags (radian/basecall.py:19-37).

The decode hot path (assemble_matrices + beam_search, basecall.py:99-123) runs on the GPU through
this package.  The signal model that turns fast5 signal into posteriors (basecall.py:60-96) is out
of scope of this build, so the posteriors are read from ``*.npz`` files in
``fast5_dir`` instead: one file per batch, holding for every read ``<read_id>`` an object array /
list of the per-window (rows, 5) float32 matrices the sig model would have produced (already
trimmed as at basecall.py:96).  Everything downstream of that point follows the reference:
global vs chunk decode, the RNA model JSON, FASTA records ``>{id}\\n{seq[::-1]}\\n`` and the
1000-reads-per-file rollover (basecall.py:129-138).

Single intentional deviation: ``--rna-model None`` switches the RNA model off.  In the reference
the string "None" is passed through and raises a TypeError inside beam_search.
"""
from __future__ import annotations

import argparse
import json
from pathlib import Path
from time import time

import numpy as np


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description="Basecall a nanopore dRNA sequencing run.")
    parser.add_argument("fast5_dir", help="Directory of fast5 files / posterior .npz batches (see module docstring).")
    parser.add_argument("fasta_dir", help="Directory to output fasta files.")
    parser.add_argument("--local", action="store_true")
    parser.add_argument("--chunk-len", default=1024, type=int)
    parser.add_argument("--step-size", default=128, type=int)
    parser.add_argument("--batch-size", default=32, type=int)
    parser.add_argument("--outlier-clip", default=4, type=int)
    parser.add_argument("--rna-model", default="models/rnamodel_12mer_pc.json")
    parser.add_argument("--sig-model", default="models/sig2seq.h5")
    parser.add_argument("--sig-config", default="models/sig2seq.yaml")
    parser.add_argument("--beam-width", default=6, type=int)
    parser.add_argument("--decode-type", choices=["global", "chunk"], default="global")
    parser.add_argument("--sig-threshold", default=0.5, type=float)
    parser.add_argument("--rna-threshold", default=0.5, type=float)
    parser.add_argument("--context-len", default=11, type=int)
    return parser


def load_rna_model(path: str, context_len: int, device: int = 0):
    """basecall.py:47-57, straight into the resident dense table."""
    from .decode import RnaTable

    if path == "None":
        return None
    table = RnaTable.from_json(path, device)
    if table.L != context_len:
        raise KeyError(f"--context-len {context_len} but {path} holds {table.L}-symbol contexts")
    return table


class FastaWriter:
    """reads-{n}.fasta with a rollover every 1000 reads (basecall.py:64-67,129-141)."""

    def __init__(self, fasta_dir: str, per_file: int = 1000):
        self.dir = fasta_dir
        self.per_file = per_file
        self.n = 0
        self.i = 0
        self.f = open(f"{fasta_dir}/reads-{self.n}.fasta", "w")

    def write(self, read_id: str, sequence: str):
        # reverse to 5'->3' exactly as basecall.py:129
        self.f.write(f">{read_id}\n{sequence[::-1]}\n")
        self.i += 1
        if self.i == self.per_file:
            self.f.close()
            self.n += 1
            self.f = open(f"{self.dir}/reads-{self.n}.fasta", "w")
            self.i = 0

    def close(self):
        self.f.close()


def basecall_batch(read_ids, chunk_lists, args, table):
    """The reference's per-read body (basecall.py:98-123) for a whole batch of reads."""
    from .decode import beam_search_batch
    from .matrix_assembly import assemble_batch

    if args.decode_type == "global":
        mats = assemble_batch(chunk_lists, args.step_size)
        return beam_search_batch(mats, args.beam_width, table, args.sig_threshold, args.rna_threshold,
                                 args.context_len)
    # chunk mode (basecall.py:110-123): every window decoded with the model off, then all reads'
    # fragments stitched in one call
    from .sequence_assembly import stitch_flat

    flat = [m for mats in chunk_lists for m in mats]
    sym, frag_off = beam_search_batch(flat, args.beam_width, None, None, None, None, return_symbols=True)
    ranges = np.zeros(len(chunk_lists) + 1, dtype=np.int64)
    ranges[1:] = np.cumsum([len(mats) for mats in chunk_lists])
    seq, off = stitch_flat(sym, frag_off, ranges)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    return [lut[seq[off[r]:off[r + 1]]].tobytes().decode("ascii") for r in range(len(chunk_lists))]


def load_posterior_batch(path):
    """A batch of window posteriors from a ``.npz`` file -> (read ids, list of window-matrix lists).
    Per read ``<id>`` either a ``(n_windows, rows, 5)`` array (all windows full), or a ``(total rows, 5)``
    array of all windows back to back plus ``<id>__lens`` with the row count of every window (the last
    one is trimmed, basecall.py:96).  Arrays of Python objects are pickles, and unpickling a file
    executes what is in it: they are only read with ``RADIAN_ALLOW_PICKLE=1`` in the environment."""
    import os

    allow = os.environ.get("RADIAN_ALLOW_PICKLE", "") not in ("", "0")
    z = np.load(path, allow_pickle=allow)
    read_ids, chunk_lists = [], []
    for key in z.files:
        if key.endswith("__lens"):
            continue
        try:
            arr = z[key]
        except ValueError as e:
            raise ValueError(f"{path}: '{key}' is an object array (a pickle); store (rows, 5) float32 plus "
                             f"'{key}__lens', or set RADIAN_ALLOW_PICKLE=1 for files you trust") from e
        if arr.dtype == object:
            mats = [np.asarray(m, dtype=np.float32) for m in arr]
        elif key + "__lens" in z.files:
            arr = np.asarray(arr, dtype=np.float32).reshape(-1, 5)
            ends = np.cumsum(z[key + "__lens"])
            mats = [arr[e - n:e] for e, n in zip(ends, z[key + "__lens"])]
        elif arr.ndim == 3:
            mats = list(np.asarray(arr, dtype=np.float32))
        else:
            raise ValueError(f"{path}: '{key}' needs '{key}__lens' or the shape (n_windows, rows, 5)")
        read_ids.append(key)
        chunk_lists.append(mats)
    return read_ids, chunk_lists


def windows_from_fast5(path, args):
    """basecall.py:70-83 for every read of one fast5 file: raw signal -> mad_normalise ->
    get_windows, all reads of the file in two GPU calls.  -> list of (read_id, windows, pad).
    A read whose signal the reference refuses is reported and skipped as at basecall.py:79-82."""
    from . import fast5, preprocess

    ids, sigs = [], []
    for rid, sig in fast5.reads(path):
        ids.append(rid)
        sigs.append(sig)
    norm = preprocess.mad_normalise_batch(sigs, args.outlier_clip)
    keep = []
    for rid, r in zip(ids, norm):
        if isinstance(r, ValueError):
            print(r.args)
            print(f"{rid} signal issue, skipping this read.")
        else:
            keep.append((rid, r))
    wins = preprocess.get_windows_batch([r for _, r in keep], args.chunk_len, args.step_size)
    return [(rid, w, pad) for (rid, _), (w, pad) in zip(keep, wins)]


def main(argv=None, sig_model=None):
    """``sig_model``: optional callable ``windows (n, chunk_len) -> (n, chunk_len, 5)`` posteriors
    standing in for the reference's Keras model (basecall.py:60-61, 86-93); with it ``*.fast5``
    files in ``fast5_dir`` are basecalled from the raw signal, without it only ``*.npz`` posterior
    batches are."""
    args = build_parser().parse_args(argv)
    table = load_rna_model(args.rna_model, args.context_len)
    fasta = FastaWriter(args.fasta_dir)
    for path in sorted(Path(args.fast5_dir).rglob("*.fast5")):
        if sig_model is None:
            print(f"{path.name}: no signal model given (out of scope of this build), skipping raw signal file.")
            continue
        start_t = time()
        reads = windows_from_fast5(path, args)
        chunk_lists = []
        for _, windows, pad in reads:
            mats = []
            for i in range(0, len(windows), args.batch_size):  # basecall.py:86-93
                mats.extend(np.asarray(sig_model(windows[i:i + args.batch_size]), dtype=np.float32))
            mats[-1] = mats[-1][:-pad]  # basecall.py:96
            chunk_lists.append(mats)
        read_ids = [rid for rid, _, _ in reads]
        for rid, seq in zip(read_ids, basecall_batch(read_ids, chunk_lists, args, table)):
            fasta.write(rid, seq)
        print(f"Basecalled {len(read_ids)} reads of {path.name} in {time() - start_t:.2f} sec.")
    for path in sorted(Path(args.fast5_dir).rglob("*.npz")):
        start_t = time()
        read_ids, chunk_lists = load_posterior_batch(path)
        seqs = basecall_batch(read_ids, chunk_lists, args, table)
        for rid, seq in zip(read_ids, seqs):
            fasta.write(rid, seq)
        print(f"Basecalled {len(read_ids)} reads of {path.name} in {time() - start_t:.2f} sec.")
    fasta.close()


if __name__ == "__main__":
    main()
