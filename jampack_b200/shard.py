"""Block sharding for the N-GPU runs: whole blocks are independent, so they are dealt out to the
ranks round-robin -- the order the reference's batch loop hands blocks to its workers (jampack.cpp:205-224) --
with no data-path collective. torch.distributed is only used for the barrier and for reducing the timings
(max over ranks) and byte counts (sum) that bench.py reports."""
import torch
import torch.distributed as dist


def shard_blocks(n_blocks, world, rank):
    """Indices of the blocks rank `rank` of `world` processes owns."""
    return list(range(rank, n_blocks, world))


def owner_of(block, world):
    return block % world


def reduce_step_stats(local_ms, local_bytes, device=None):
    """-> (max over ranks of local_ms, sum over ranks of local_bytes). Identity without a process group."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(local_ms), int(local_bytes)
    dev = device if device is not None else torch.device("cpu")
    t = torch.tensor([float(local_ms)], dtype=torch.float64, device=dev)
    b = torch.tensor([int(local_bytes)], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(b, op=dist.ReduceOp.SUM)
    return float(t.item()), int(b.item())


def gather_digests(local, device=None):
    """All ranks' {block index: digest} dicts merged on every rank (used to check a sharded run end to end)."""
    if not (dist.is_available() and dist.is_initialized()):
        return dict(local)
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, dict(local))
    merged = {}
    for d in out:
        merged.update(d)
    return merged
