"""N>1 path on CPU: two gloo ranks shard a batch (LPT), each decodes its shard, rank 0 gathers the
sequences in read order.  The per-rank decoder is the CPU oracle here (no GPU in this suite); on
the GPU box the same code path runs with the CUDA decoder (tests/test_gpu_parity.py covers it)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def oracle_decode(mats, beam_width, lm, s_thr, r_thr, L):
    from oracle import oracle

    return ["".join("ACGT"[s] for s in oracle.beam_search(m, beam_width, lm, L or 0, s_thr, r_thr)[0]) for m in mats]


def make_inputs():
    from radian_b200 import synth

    post, off = synth.make_reads(np.array([12, 40, 7, 25, 33, 18, 9]), seed=2)
    post = post.numpy()
    off = off.numpy()
    return [post[off[i]:off[i + 1]] for i in range(len(off) - 1)], synth.make_table(3, 4)


def worker(rank, world, port, q):
    import sys

    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from radian_b200 import parallel

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    mats, tab = make_inputs()
    out = parallel.decode_sharded(mats, 6, tab, 0.5, 0.5, 3, decode_fn=oracle_decode)
    mine = parallel.shard_for_rank([len(m) for m in mats], rank, world)
    q.put((rank, out, mine.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_decode():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        r, out, mine = q.get(timeout=120)
        res[r] = (out, mine)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    mats, tab = make_inputs()
    want = oracle_decode(mats, 6, tab, 0.5, 0.5, 3)
    assert res[0][0] == want          # rank 0 holds everything, in read order
    assert res[1][0] is None          # nothing is replicated to the other rank
    assert sorted(res[0][1] + res[1][1]) == list(range(len(mats)))   # disjoint, complete shards
