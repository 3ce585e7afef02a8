"""Multi-GPU layer: independent replicas (temperatures / restarts / copies) are sharded over ranks, one process per GPU.

The reference's only parallelism is a Slurm job farm with one process per temperature (src/Launcher.jl:239-345) whose results
are merged from files (src/IO.jl:64-143).  Here a rank owns a contiguous range of GLOBAL replica ids; the ids enter the
counter-based random stream, so every replica's chain is bit-identical however the replicas are sharded (tests/test_distributed.py).
There is no data-path collective; the only communication is the final gather of per-replica observables
(torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_range(n_total, world, rank):
    """Contiguous [offset, offset + count) of `n_total` replicas owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n_total), int(world))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def shard(array, world, rank, axis=0):
    off, cnt = shard_range(np.shape(array)[axis], world, rank)
    return np.take(np.asarray(array), np.arange(off, off + cnt), axis=axis)


def gather_replicas(local, n_total, group=None, device=None):
    """All-gather per-replica rows (first axis = this rank's replicas, in global order) into the full [n_total, ...] array.
    Works with any initialised torch.distributed backend; without one (single process) it returns `local` unchanged."""
    import torch
    import torch.distributed as dist
    local = np.ascontiguousarray(local)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        assert local.shape[0] == n_total
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = -(-int(n_total) // world)                                   # padded shard size: all_gather needs equal shapes
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    if local.shape[0] == per:
        t = torch.from_numpy(local).to(device)
    else:
        buf = np.zeros((per,) + local.shape[1:], dtype=local.dtype)
        buf[:local.shape[0]] = local
        t = torch.from_numpy(buf).to(device)
    # one collective into one tensor and ONE copy back to the host (a list of per-rank tensors costs a copy and a synchronisation per rank)
    full = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=t.dtype, device=device)
    try:
        dist.all_gather_into_tensor(full, t, group=group)
    except (RuntimeError, NotImplementedError, AttributeError):      # backend without the single-tensor form
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t, group=group)
        full = torch.cat(out, dim=0)
    host = full.cpu().numpy()
    if n_total == world * per:
        return host
    parts = []
    for r in range(world):
        _, cnt = shard_range(n_total, world, r)
        parts.append(host[r * per:r * per + cnt])
    return np.concatenate(parts, axis=0)


def reduce_min_energy(E_local, n_total, group=None):
    """Best restart over all ranks: (global replica id, energy) of the minimum (annealing restarts, BASELINE config C4)."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    E_all = gather_replicas(np.asarray(E_local, dtype=np.float64), n_total, group)
    k = int(np.argmin(E_all))
    return k, float(E_all[k])
