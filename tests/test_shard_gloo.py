"""N>1 host logic on CPU: two gloo ranks shard a corpus of blocks, run the stage (here the ORACLE stands in for
the GPU stage -- this test is about the sharding/reduction plumbing, not the kernels) and must together
reproduce the single-process result."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_blocks, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import oracle
    from jampack_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.shard_blocks(n_blocks, world, rank)
    digests, nbytes = {}, 0
    for b in mine:
        T = oracle.gen("markov2", 120 * 200 + b, 100 + b)
        out = oracle.forward(T, "port")
        assert (oracle.inverse(out, "port") == T).all()
        digests[b] = oracle.fnv(out)
        nbytes += T.size
    dist.barrier()
    ms, total = shard.reduce_step_stats(10.0 * (rank + 1), nbytes)
    merged = shard.gather_digests(digests)
    q.put((rank, mine, ms, total, merged))
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    sys.path.insert(0, ROOT)
    import oracle
    from jampack_b200 import shard
    n_blocks, world = 7, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_blocks, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(world)]
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    owned = sorted(b for _, mine, _, _, _ in res for b in mine)
    assert owned == list(range(n_blocks))                               # a partition: no block twice, none missing
    want = {}
    total = 0
    for b in range(n_blocks):
        T = oracle.gen("markov2", 120 * 200 + b, 100 + b)
        want[b] = oracle.fnv(oracle.forward(T, "port"))
        total += T.size
        assert shard.owner_of(b, world) == b % world
    for rank, mine, ms, tot, merged in res:
        assert ms == 20.0                                               # max over ranks
        assert tot == total                                             # sum over ranks
        assert merged == want                                           # every rank sees the whole job's digests


def test_single_process_is_identity():
    sys.path.insert(0, ROOT)
    from jampack_b200 import shard
    assert shard.shard_blocks(5, 1, 0) == [0, 1, 2, 3, 4]
    assert shard.reduce_step_stats(3.5, 100) == (3.5, 100)
    assert shard.gather_digests({1: 2}) == {1: 2}
