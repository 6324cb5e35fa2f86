"""World-size-2 gloo test (CPU) of the multi-GPU host logic: frame pairs / camera streams are
sharded round-robin with no data-path collective, and the per-unit scalars are gathered."""
import os
import socket

import numpy as np
import pytest

from opticalflowcontainer_b200 import sharding


def test_shard_indices_partition():
    for n in (0, 1, 7, 8, 64):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                idx = sharding.shard_indices(n, r, world)
                assert all(sharding.owner_of(i, world) == r for i in idx)
                seen += idx
            assert sorted(seen) == list(range(n))
    assert sharding.split_batches(list(range(7)), 3) == [[0, 1, 2], [3, 4, 5], [6]]
    with pytest.raises(ValueError):
        sharding.shard_indices(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_units, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.shard_indices(n_units, rank, world)
    # stand-in for the per-pair engine call: a deterministic function of the unit index
    local = {i: float(i) * 0.5 - 3.0 for i in mine}
    out = sharding.gather_unit_values(local, n_units)
    dist.barrier()
    q.put((rank, mine, out))
    dist.destroy_process_group()


def test_gather_over_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port, n_units, world = _free_port(), 11, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_units, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.arange(n_units) * 0.5 - 3.0
    owned = []
    for rank, mine, out in res:
        assert np.allclose(out, expect)
        owned += mine
    assert sorted(owned) == list(range(n_units))


def test_gather_without_process_group():
    out = sharding.gather_unit_values({0: 1.5, 2: -2.0}, 4)
    assert out[0] == 1.5 and out[2] == -2.0 and np.isnan(out[1]) and np.isnan(out[3])
