"""Read sharding over the GPUs of one box with the CUDA decoder: every rank decodes its
LPT shard on its own GPU against its own table replica, rank 0 gathers the strings on the host
(`parallel.decode_sharded`), and the result equals the oracle's, read by read.  Needs two GPUs
(`gpurun --gpus 2`); skipped on a one-GPU box.  The host-side logic alone runs on CPU in
tests/test_parallel_gloo.py."""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys, types
    sys.path.insert(0, %(root)r)
    import numpy as np
    import torch
    import torch.distributed as dist
    from radian_b200 import basecall, decode, parallel, synth

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")  # the only exchange is the host-side gather of strings
    nb = synth.read_lengths(48, 7, median=300, lo=20, hi=3000)
    post, off = synth.make_reads(nb, seed=5)
    post, off = post.numpy(), off.numpy()
    mats = [post[off[i]:off[i + 1]] for i in range(len(nb))]
    tab_np = synth.make_table(9, 5)
    table = decode.RnaTable(tab_np, local)
    # configs[2]: RNA-LM global decode, bw 16
    seqs = parallel.decode_sharded(mats, 16, table, 0.5, 0.5, 9)
    # configs[3] chunk mode: every rank's shard through basecall_batch (decode of all windows + stitch)
    args = types.SimpleNamespace(decode_type="chunk", beam_width=16, step_size=128)
    def chunk_decode(sub, bw, lm, s, r, L):
        return basecall.basecall_batch(None, [synth.split_windows(m, 1024, 128) for m in sub], args, None)
    cseqs = parallel.decode_sharded(mats, 16, None, None, None, None, decode_fn=chunk_decode)
    if rank == 0:
        from oracle import oracle
        assert len(seqs) == len(mats) and len(cseqs) == len(mats)
        for i in (0, 5, 17, 30, 47):
            want = "".join("ACGT"[s] for s in oracle.beam_search(mats[i], 16, tab_np, 9, 0.5, 0.5, topk=1)[0])
            assert seqs[i] == want, i
            frags = ["".join("ACGT"[s] for s in oracle.beam_search(w, 16, topk=1)[0]) for w in synth.split_windows(mats[i], 1024, 128)]
            assert cseqs[i] == oracle.stitch(frags)[0], i
        mine = parallel.shard_for_rank([len(m) for m in mats], 0, world)
        assert 0 < len(mine) < len(mats)
        print("SHARDED-OK", world, sum(len(s) for s in seqs))
    else:
        assert seqs is None and cseqs is None
    dist.destroy_process_group()
''')


def test_decode_sharded_on_two_gpus(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "SHARDED-OK 2" in r.stdout
