"""world_size-2 gloo tests (CPU) of the N>1 host path: balanced clip sharding and the detection all-gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from class_query_vad_b200.dist import shard_range, gather_detections, pack_detections


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 32, 33, 255):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_clips, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(1234)
        full = torch.randn(n_clips, 3, 11, generator=g)            # the "single-GPU" result, identical on all ranks
        lo, hi = shard_range(n_clips, world, rank)
        logits, boxes, lb = full[lo:hi, :, :4], full[lo:hi, :, 4:8], full[lo:hi, :, 8:]
        det = pack_detections(logits, boxes, lb)
        out_a = gather_detections(det, n_clips_total=n_clips)
        out_b = gather_detections(det)                              # sizes exchanged by a first all-gather
        ok = bool(torch.equal(out_a, full) and torch.equal(out_b, full))
        if rank == 0:
            q.put(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [8, 7])      # even and ragged shards
def test_gather_detections_world2_gloo(n_clips):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_clips, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get() is True


def test_gather_is_identity_without_process_group():
    x = torch.randn(3, 2, 5)
    assert gather_detections(x) is x


def _allreduce_worker(rank, world, port, q):
    import os
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from class_query_vad_b200.dist import allreduce_gradients

    class FakeEngine:      # the engine contract allreduce_gradients relies on: one flat fp32 gradient buffer
        _gflat = torch.arange(10, dtype=torch.float32) * (rank + 1)
    out = allreduce_gradients(FakeEngine())
    q.put((rank, out.tolist()))
    dist.destroy_process_group()


def test_gradient_allreduce_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650
    procs = [ctx.Process(target=_allreduce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    expect = [1.5 * i for i in range(10)]      # mean of (1x, 2x)
    assert res[0] == expect and res[1] == expect


def test_gradient_bucket_layout_is_one_contiguous_bucket_per_layer():
    """Host logic of the overlapped gradient all-reduce: the flat gradient buffer holds layers.l.* + cls_layers.l.* contiguously per
    layer (bucket l), the shared modules last; buckets tile the buffer without gaps or overlap."""
    from class_query_vad_b200 import _lib
    from class_query_vad_b200.engine import grad_bucket_layout
    lib = _lib.lib()
    layers = 6
    n = lib.cqvad_decoder_num_weights(layers)
    names = [lib.cqvad_decoder_weight_name(i, layers).decode() for i in range(n)]
    sizes = [(i * 37) % 1000 + 1 for i in range(n)]
    offs, buckets = grad_bucket_layout(names, sizes, layers)
    assert len(buckets) == layers + 1 and buckets[0][0] == 0 and buckets[-1][1] == offs[-1]
    for (lo, hi), (lo2, _) in zip(buckets, buckets[1:]):
        assert hi == lo2 and hi > lo
    for i, name in enumerate(names):
        parts = name.split(".")
        b = int(parts[1]) if parts[0] in ("layers", "cls_layers") else layers
        lo, hi = buckets[b]
        assert lo <= offs[i] and offs[i] + sizes[i] <= hi, name
    spans = sorted((int(offs[i]), int(offs[i]) + sizes[i]) for i in range(n))
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))          # no two gradients overlap


def _wire_worker(rank, world, port, n_clips, out_dir, q):
    """Evaluation loop of the N>1 path end to end on the host side: shard -> per-rank detections -> all-gather -> rank 0 writes the
    reference's text file and the binary dump (class_query_vad_b200/detections.py)."""
    import numpy as np
    from class_query_vad_b200.detections import save_detections, write_reference_text
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        K, nq = 6, 3
        g = torch.Generator().manual_seed(99)
        full = torch.rand(n_clips, nq, K + 5, generator=g)          # [scores | boxes | person] of every clip
        lo, hi = shard_range(n_clips, world, rank)
        det = gather_detections(full[lo:hi].contiguous(), n_clips_total=n_clips)
        if rank == 0:
            ids = [f"v{b},{b:04d}" for b in range(n_clips)]
            write_reference_text(os.path.join(out_dir, "0.txt"), ids, det, K)
            save_detections(os.path.join(out_dir, "det.npz"), ids, det, K)
            q.put(bool(torch.equal(det, full)))
    finally:
        dist.destroy_process_group()


def test_detection_wire_format_world2_gloo(tmp_path):
    import numpy as np
    from class_query_vad_b200.detections import load_detections, read_reference_text
    n_clips, K, nq = 5, 6, 3
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_wire_worker, args=(r, 2, port, n_clips, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get() is True
    full = torch.rand(n_clips, nq, K + 5, generator=torch.Generator().manual_seed(99)).numpy()
    ids, det = load_detections(tmp_path / "det.npz")
    assert ids == [f"v{b},{b:04d}" for b in range(n_clips)] and np.array_equal(det, full)
    lid, boxes, scores, person = read_reference_text(tmp_path / "0.txt", K)
    flat = full.reshape(-1, K + 5).astype(np.float64)
    assert lid == [i for i in ids for _ in range(nq)]
    assert np.array_equal(boxes, flat[:, K:K + 4]) and np.array_equal(scores, flat[:, :K]) and np.array_equal(person, flat[:, K + 4])
