"""CPU, world_size 2 over gloo: the N>1 host path -- shard bookkeeping and the one-collective packed gather of the wire
records (sharding.PackedGather), with even and uneven shards.  The records are real ones (wire.pack of people lists), and
what rank 0 ends up with must unpack to exactly the per-image lists of a single-process run, in image order."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _people_for(image: int):
    """Deterministic fake people of one image: `image % 4` persons, joint g of person p at (image + g/32, p + 0.5)."""
    out = []
    for p in range(image % 4):
        pts = [((0, 0) if (g + p + image) % 5 == 0 else (np.float64(image + g / 32.0), np.float64(p + 0.5))) for g in range(17)]
        out.append((pts, np.float64(1.0 - 1.0 / (image + p + 2))))
    return out


def _worker(rank, world, port, n_images, q):
    import torch
    import torch.distributed as dist

    from improved_body_parts_b200 import wire
    from improved_body_parts_b200.sharding import PackedGather, shard_range

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n_images, rank, world)
        rb = wire.record_bytes(17, 4)
        pg = PackedGather(hi - lo, rb, device="cpu", dst=0)
        assert pg.counts == [shard_range(n_images, r, world)[1] - shard_range(n_images, r, world)[0] for r in range(world)]
        out_ptr = pg.out.data_ptr() if pg.out is not None else None
        for step in range(3):  # the same pre-allocated buffers every step
            rec = wire.pack([_people_for(i + 100 * step) for i in range(lo, hi)], rows=4, status=[i for i in range(lo, hi)])
            pg.local[:hi - lo].copy_(torch.from_numpy(rec.view(np.uint8).reshape(hi - lo, rb)))
            pg.gather()
            assert (pg.out.data_ptr() if pg.out is not None else None) == out_ptr
        got = pg.records()
        if rank == 0:
            q.put(got.numpy().copy())
        else:
            assert got is None
            q.put("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [12, 13, 1])
def test_packed_gather_returns_image_order_on_rank0(n):
    import torch.multiprocessing as mp

    from improved_body_parts_b200 import wire

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got = next(r for r in results if not isinstance(r, str))
    rec = wire.as_records(got, 17, 4)
    assert len(rec) == n
    assert list(rec["status"]) == list(range(n))  # rank order == image order, padding of the short shard trimmed
    back = wire.unpack(rec, list(range(n)))
    for i in range(n):
        want = _people_for(i + 200)  # the last step's records
        assert len(back[i]) == len(want)
        for (pa, sa), (pb, sb) in zip(want, back[i]):
            assert sa == sb and all(a[0] == b[0] and a[1] == b[1] and isinstance(a[0], int) == isinstance(b[0], int)
                                    for a, b in zip(pa, pb))


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
