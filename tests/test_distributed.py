"""N > 1 host logic on CPU: world_size-2 gloo processes shard a replica set and gather per-replica rows.
The device chains themselves are sharding-invariant by construction (tests/test_gpu_parity.py::test_replica_sharding_is_invisible)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition(scmc):
    from scmc_b200.distributed import shard_range
    for n, w in ((8192, 8), (4096, 8), (64, 3), (5, 8), (1, 1)):
        spans = [shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        for (o1, c1), (o2, _) in zip(spans, spans[1:]):
            assert o1 + c1 == o2
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import scmc_b200  # noqa: F401
    from scmc_b200.distributed import shard_range, gather_replicas, reduce_min_energy
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    off, cnt = shard_range(n_total, world, rank)
    gid = np.arange(off, off + cnt)
    rows = np.stack([gid * 1.0, np.sin(gid)], axis=1)            # stand-in for per-replica observable rows
    full = gather_replicas(rows, n_total)
    k, e = reduce_min_energy(np.cos(gid * 0.7), n_total)
    q.put((rank, full, k, e))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 64])
def test_gloo_gather_world2(n_total):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    gid = np.arange(n_total)
    expect = np.stack([gid * 1.0, np.sin(gid)], axis=1)
    for rank, full, k, e in res:
        assert np.array_equal(full, expect)
        assert k == int(np.argmin(np.cos(gid * 0.7))) and np.isclose(e, np.cos(gid * 0.7).min())
