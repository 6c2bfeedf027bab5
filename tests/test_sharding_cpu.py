"""Host logic of the N>1 path on CPU: contiguous image sharding and the single all-gather of padded
detections, with world_size 2 over gloo (the GPU path uses the same functions over NCCL)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sodt_b200.runtime import DetectionBuffer, ShardedDetector, allgather_detections, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_detections(image_index, max_det=300):
    """Deterministic detections of a global image index (what NMS would have written)."""
    g = torch.Generator().manual_seed(1000 + image_index)
    n = int(torch.randint(0, 40, (1,), generator=g))
    det = torch.zeros(max_det, 6)
    det[:n] = torch.rand(n, 6, generator=g)
    return det, n


def _worker(rank, world, port, n_images, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n_images, rank, world)
        buf = DetectionBuffer(hi - lo, "cpu")
        for j, idx in enumerate(range(lo, hi)):
            d, n = _fake_detections(idx)
            buf.det[j].copy_(d)
            buf.counts[j] = n
        det, counts = allgather_detections(buf)
        ok = det.shape == (n_images, 300, 6) and counts.shape == (n_images,)
        for idx in range(n_images):
            d, n = _fake_detections(idx)
            ok = ok and int(counts[idx]) == n and torch.equal(det[idx], d)
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_contiguously():
    for n, world in [(64, 8), (10, 4), (3, 8), (32, 1), (7, 2)]:
        spans = [shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_detection_buffer_views_alias_flat():
    buf = DetectionBuffer(3, "cpu")
    buf.det[1, 5, 2] = 7.5
    buf.counts[2] = 123
    d, c = DetectionBuffer.split(buf.flat.clone(), 3)
    assert d[1, 5, 2] == 7.5 and int(c[2]) == 123 and int(c[0]) == 0
    assert buf.flat.numel() == 3 * 300 * 6 + 3


def test_allgather_detections_world2_gloo():
    world, n_images = 2, 8
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_images, results), nprocs=world, join=True)
    assert dict(results) == {0: True, 1: True}


class _FakeDetector:
    """Stands in for runtime.Detector on the CPU: 'detects' the deterministic boxes of the global image indices it is handed."""
    device = torch.device("cpu")

    def detect_device(self, indices, _ir):
        buf = DetectionBuffer(len(indices), "cpu")
        for j, idx in enumerate(indices):
            d, n = _fake_detections(int(idx))
            buf.det[j].copy_(d)
            buf.counts[j] = n
        return buf


def _sharded_worker(rank, world, port, n_images, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = ShardedDetector(_FakeDetector())
        ok = sh.rank == rank and sh.world == world and sh._pending is None and sh._step == 0
        lo, hi = sh.local_slice(n_images)
        for step in range(2):                                   # two steps: the per-step state (receive ring, step counter) advances
            g = sh.detect_device(list(range(lo, hi)), None)
            det, counts = g.tensors()
            ok = ok and det.shape == (n_images, 300, 6) and sh._step == step + 1
            for idx in range(n_images):
                d, n = _fake_detections(idx)
                ok = ok and int(counts[idx]) == n and torch.equal(det[idx], d)
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_sharded_detector_host_logic_world2_gloo():
    """ShardedDetector end to end on the CPU (constructor state, local_slice, detect_device's gather, GatheredDetections views)."""
    world, n_images = 2, 6
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_sharded_worker, args=(world, port, n_images, results), nprocs=world, join=True)
    assert dict(results) == {0: True, 1: True}
