"""Read sharding across GPUs: one process per GPU, index replicated, read pairs split into
the reference's own batches (SOAP4.cpp:206: 2,097,152 reads = 1,048,576 pairs) and dealt round-robin to the
ranks.  There is no collective on the data path; torch.distributed only carries
  * the first-batch parameters every shard must share (detected read lengths, insert_low clamp,
    SOAP4.cpp:458-521) -- broadcast from rank 0,
  * the end-of-run counters (pairs aligned / alignments / unaligned) -- summed,
  * the per-shard output streams' order -- shard outputs are concatenated in batch order.
Works with the gloo backend on CPU (tests) and nccl on GPUs (bench.py)."""
import torch
import torch.distributed as dist

BATCH_PAIRS = 12 * 8192 * 128 // 6 // 2          # maxNumQueries / 2 (SOAP4.cpp:206)


def batches(total_pairs, batch_pairs=BATCH_PAIRS):
    """[(first_pair, n_pairs)] in input order."""
    out, p = [], 0
    while p < total_pairs:
        n = min(batch_pairs, total_pairs - p)
        out.append((p, n))
        p += n
    return out


def my_batches(total_pairs, rank, world, batch_pairs=BATCH_PAIRS):
    """Batches of this rank: batch b goes to rank b % world."""
    return [(b, first, n) for b, (first, n) in enumerate(batches(total_pairs, batch_pairs)) if b % world == rank]


def first_batch_params(read_lengths_first_batch, insert_low, device="cpu"):
    """Rank 0 inspects the first batch like GetReadLength (QueryParser.cpp:2253-2277: maximum over the first 999999
    sampled reads of each mate) and clamps insert_low; every rank receives the same three numbers."""
    t = torch.zeros(3, dtype=torch.int64, device=device)
    if not dist.is_initialized() or dist.get_rank() == 0:
        l1 = int(max(read_lengths_first_batch[0::2][:999999])) if len(read_lengths_first_batch) else 100
        l2 = int(max(read_lengths_first_batch[1::2][:999999])) if len(read_lengths_first_batch) > 1 else 100
        t[0], t[1], t[2] = l1, l2, max(int(insert_low), l1, l2)
    if dist.is_initialized():
        dist.broadcast(t, src=0)
    return int(t[0]), int(t[1]), int(t[2])


def sum_counters(counters, device="cpu"):
    """counters: dict name -> int; summed over ranks (the reference prints the totals, SOAP4.cpp:599-613)."""
    keys = sorted(counters)
    t = torch.tensor([int(counters[k]) for k in keys], dtype=torch.int64, device=device)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: int(v) for k, v in zip(keys, t.tolist())}


def max_over_ranks(values, device="cpu"):
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def gather_outputs(chunks):
    """chunks: list of (batch_index, bytes) produced by this rank.  Rank 0 returns all shards' chunks
    concatenated in batch order (the single-process output order); other ranks return None."""
    if not dist.is_initialized():
        return b"".join(c for _, c in sorted(chunks))
    world, rank = dist.get_world_size(), dist.get_rank()
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(chunks, gathered, dst=0)
    if rank != 0:
        return None
    allc = [c for part in gathered for c in part]
    return b"".join(c for _, c in sorted(allc))
