"""CPU suite: the multi-GPU host logic (megapath_b200/shard.py) with world_size 2 over gloo."""
import hashlib
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp_spawn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def fake_align(first, n):
    """stand-in for one batch's output stream and counters (host logic only, no GPU)."""
    out = b"".join(b"@p%d\tSCORE:%d;\n" % (p, p % 7) for p in range(first, first + n))
    return out, {"pairs_aligned": sum(1 for p in range(first, first + n) if p % 7), "alignments": n}


def worker(rank, world, port, total, batch_pairs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from megapath_b200 import shard
    lens = np.array([150, 140, 151, 149] * 8, dtype=np.uint32) if rank == 0 else np.zeros(0, np.uint32)   # only rank 0 has seen the first batch
    l1, l2, ilow = shard.first_batch_params(lens, 1)
    chunks, cnt = [], {"pairs_aligned": 0, "alignments": 0}
    for b, first, n in shard.my_batches(total, rank, world, batch_pairs):
        out, c = fake_align(first, n)
        chunks.append((b, out))
        for k in cnt:
            cnt[k] += c[k]
    tot = shard.sum_counters(cnt)
    mx = shard.max_over_ranks([float(rank + 1)])
    merged = shard.gather_outputs(chunks)
    q.put((rank, (l1, l2, ilow), tot, mx, hashlib.md5(merged).hexdigest() if merged is not None else None))
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    from megapath_b200 import shard
    total, bp = 10_000, 1536
    assert sum(n for _, n in shard.batches(total, bp)) == total
    ctx = mp_spawn.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=worker, args=(r, 2, port, total, bp, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single_out, single_cnt = fake_align(0, total)
    (r0, par0, tot0, mx0, md0), (r1, par1, tot1, mx1, md1) = res
    assert par0 == par1 == (151, 149, 151)                 # broadcast from rank 0
    assert tot0 == tot1 == single_cnt                      # summed counters
    assert mx0 == mx1 == [2.0]                             # max over ranks
    assert md0 == hashlib.md5(single_out).hexdigest() and md1 is None    # shard outputs concatenated in batch order
    # every pair is owned by exactly one rank
    owned = sorted(p for r in range(2) for _, f, n in shard.my_batches(total, r, 2, bp) for p in (f, f + n))
    assert owned[0] == 0 and owned[-1] == total
