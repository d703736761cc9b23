"""CPU, world_size 2 over gloo: the N>1 host path -- shard bookkeeping and the gather of person lists."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_images, q):
    import torch
    import torch.distributed as dist

    from improved_body_parts_b200.sharding import gather_people, shard_range

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n_images, rank, world)
        idx = torch.arange(lo, hi)
        local = {"n_persons": (idx % 7).to(torch.int32),
                 "people_xy": idx.to(torch.float64)[:, None, None, None].expand(hi - lo, 3, 17, 2).contiguous() + 0.5,
                 "people_score": idx.to(torch.float64)[:, None].expand(hi - lo, 3).contiguous() * 2.0}
        got = gather_people(local, dst=0)
        if rank == 0:
            q.put({k: v.numpy() for k, v in got.items()})
        else:
            assert got is None
            q.put("ok")
    finally:
        dist.destroy_process_group()


def test_gather_returns_image_order_on_rank0():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port, n = _free_port(), 12
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got = next(r for r in results if isinstance(r, dict))
    idx = np.arange(n)
    assert np.array_equal(got["n_persons"], (idx % 7).astype(np.int32))
    assert np.array_equal(got["people_xy"][:, 0, 0, 0], idx + 0.5)
    assert np.array_equal(got["people_score"][:, 2], idx * 2.0)


def test_shard_ranges_partition_the_batch():
    from improved_body_parts_b200.sharding import shard_range

    for n in (0, 1, 7, 256, 2048, 2051):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)
