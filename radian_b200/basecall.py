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
    parser.add_argument("fast5_dir", help="Directory of posterior .npz batches (see module docstring).")
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
    from .sequence_assembly import stitch_batch

    flat = [m for mats in chunk_lists for m in mats]
    frags = beam_search_batch(flat, args.beam_width, None, None, None, None)
    per_read, k = [], 0
    for mats in chunk_lists:
        per_read.append(frags[k:k + len(mats)])
        k += len(mats)
    return stitch_batch(per_read)


def main(argv=None):
    args = build_parser().parse_args(argv)
    table = load_rna_model(args.rna_model, args.context_len)
    fasta = FastaWriter(args.fasta_dir)
    for path in sorted(Path(args.fast5_dir).rglob("*.npz")):
        start_t = time()
        z = np.load(path, allow_pickle=True)
        read_ids = list(z.files)
        chunk_lists = [[np.asarray(m, dtype=np.float32) for m in z[r]] for r in read_ids]
        seqs = basecall_batch(read_ids, chunk_lists, args, table)
        for rid, seq in zip(read_ids, seqs):
            fasta.write(rid, seq)
        print(f"Basecalled {len(read_ids)} reads of {path.name} in {time() - start_t:.2f} sec.")
    fasta.close()


if __name__ == "__main__":
    main()
