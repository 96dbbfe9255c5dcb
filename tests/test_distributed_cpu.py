"""World-size-2 gloo test (CPU) of the multi-GPU host logic: batch sharding by image and the final detection gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_detections(lo, hi, mx=7):
    """Deterministic per-image records so that every rank can rebuild the global answer."""
    B = hi - lo
    g = torch.arange(lo, hi, dtype=torch.float32)
    boxes = g.view(B, 1, 1).expand(B, mx, 4) * 0.001 + torch.arange(mx, dtype=torch.float32).view(1, mx, 1) * 0.01
    classes = (torch.arange(lo, hi).view(B, 1) * 3 + torch.arange(mx).view(1, mx)) % 80
    scores = 1.0 / (1.0 + g.view(B, 1) + torch.arange(mx, dtype=torch.float32).view(1, mx))
    nv = (torch.arange(lo, hi) % (mx + 1)).to(torch.int32)
    return boxes.contiguous(), classes.to(torch.int64), scores.contiguous(), nv


def _worker(rank, world, port, global_batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from yolo_v3_tf2_b200 import distributed as y3dist
    lo, hi = y3dist.shard_range(global_batch, rank, world)
    out = y3dist.gather_detections(*_fake_detections(lo, hi))
    ref = _fake_detections(0, global_batch)
    ok = all(torch.equal(a, b) for a, b in zip(out, ref))
    # the packed variant (what the graphed serving step sends): one collective on pre-packed records
    rec = y3dist.pack_detections(*_fake_detections(lo, hi))
    out2 = y3dist.unpack_detections(y3dist.gather_packed(rec))
    ok = ok and all(torch.equal(a, b) for a, b in zip(out2, ref))
    q.put((rank, lo, hi, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 12, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, 0, 6, True), (1, 6, 12, True)]


def test_shard_range_and_pack_roundtrip():
    from yolo_v3_tf2_b200 import distributed as y3dist
    assert [y3dist.shard_range(256, r, 8) for r in (0, 3, 7)] == [(0, 32), (96, 128), (224, 256)]
    with pytest.raises(ValueError):
        y3dist.shard_range(10, 0, 4)
    d = _fake_detections(0, 5)
    back = y3dist.unpack_detections(y3dist.pack_detections(*d))
    for a, b in zip(d, back):
        assert torch.equal(a, b)
    # single process: gather is the identity
    same = y3dist.gather_detections(*d)
    assert all(x is y for x, y in zip(same, d))
    rec = y3dist.pack_detections(*d)
    assert y3dist.gather_packed(rec) is rec
