"""Multi-GPU plumbing: tiles shard by index across ranks (one process per GPU, weights replicated
once), and the only exchange is the end-of-batch gather of the variable-length instance tables to
every rank -- the B200 replacement for the reference's single-process `torch.nn.DataParallel`
scatter/gather (reference infer/base.py:69, run_infer.py:139).  Works on any torch.distributed
backend (NCCL on the GPUs; gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous [lo, hi) slice of n_items owned by `rank` (first ranks take the remainder)."""
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_tables(table, nrows, group=None):
    """table [B,max_rows,10] int64, nrows [B] int32 (same B and max_rows on every rank) ->
    (table_all [world*B,max_rows,10], nrows_all [world*B]) on every rank, rank-major order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return table, nrows
    t_all = torch.empty((world * table.shape[0],) + tuple(table.shape[1:]), dtype=table.dtype, device=table.device)
    n_all = torch.empty((world * nrows.shape[0],), dtype=nrows.dtype, device=nrows.device)
    dist.all_gather_into_tensor(n_all, nrows.contiguous(), group=group)
    dist.all_gather_into_tensor(t_all, table.contiguous(), group=group)
    return t_all, n_all


def compact_rows(table_all, nrows_all):
    """Drop the padding: list of [n_i,10] arrays, one per tile, in global tile order."""
    t = table_all.cpu().numpy()
    n = nrows_all.cpu().numpy()
    return [t[i, : int(n[i])] for i in range(t.shape[0])]
