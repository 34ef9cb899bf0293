"""Multi-GPU plumbing of the decoder path: one process per GPU, clips sharded across ranks, no data-path collective
inside the decoder (clips are independent: SURVEY.md section 8e); one all-gather of the per-clip detections replaces the
reference's per-rank text files + barrier (utils/video_action_recognition.py:231-255).

NCCL over NVLink on GPUs; the same code runs on the `gloo` backend for the CPU tests (tests/test_dist_cpu.py).
"""
import torch
import torch.distributed as dist


def shard_range(n_clips, world, rank):
    """Contiguous, balanced shard [lo, hi) of n_clips for `rank` (first n_clips % world ranks get one extra clip)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(n_clips, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_detections(pred_logits, pred_boxes, pred_logits_b):
    """[B, nq, K | 4 | 3] rows: class scores, box, person logits -- the row format the reference writes per detection
    (utils/video_action_recognition.py:234)."""
    return torch.cat([pred_logits, pred_boxes, pred_logits_b], dim=-1).contiguous()


def gather_detections(det_local, n_clips_total=None, group=None):
    """All-gather per-rank detections [B_local, nq, D] into [sum B_local, nq, D] in rank order.  Ragged shards (as
    produced by shard_range) are padded to the largest shard for the collective and trimmed afterwards."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return det_local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if n_clips_total is None:
        sizes = torch.tensor([det_local.shape[0]], dtype=torch.int64, device=det_local.device)
        all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
        dist.all_gather(all_sizes, sizes, group=group)
        counts = [int(s.item()) for s in all_sizes]
    else:
        counts = [shard_range(n_clips_total, world, r)[1] - shard_range(n_clips_total, world, r)[0] for r in range(world)]
    if counts[rank] != det_local.shape[0]:
        raise ValueError(f"rank {rank}: local shard has {det_local.shape[0]} clips, expected {counts[rank]}")
    bmax = max(counts)
    padded = det_local
    if det_local.shape[0] != bmax:
        padded = det_local.new_zeros((bmax,) + tuple(det_local.shape[1:]))
        padded[: det_local.shape[0]] = det_local
    out = det_local.new_empty((world * bmax,) + tuple(det_local.shape[1:]))
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    if all(c == bmax for c in counts):
        return out
    return torch.cat([out[r * bmax: r * bmax + counts[r]] for r in range(world)], dim=0)


def allreduce_gradients(engine, group=None, average=True, overlap=False, comm_stream=None):
    """Training: all-reduce of the engine's flat fp32 gradient buffer (all decoder parameters, ~133 MB for the 6-layer AVA decoder)
    once per optimizer step -- what DDP does per bucket in the reference (utils/model_utils.py:113-121).  Parameters the reference
    never uses (q_proj, cls_norm) are not in the buffer, which is what static_graph=True DDP skips.

    overlap=False: ONE collective after the backward.  overlap=True (call it right after `engine.backward(...)` returned, i.e.
    while the GPU is still executing that backward; needs `engine.enable_layer_events()`): one collective per decoder layer on
    `comm_stream`, each waiting only for its layer's gradient-complete event, layers-1 first, then the shared-module bucket after
    the whole backward; the caller's stream waits for the communication stream at the end.  The average is folded into the
    buckets as they complete."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return engine._gflat
    world = dist.get_world_size(group)
    if not overlap or getattr(engine, "layer_events", None) is None:
        _reduce_bucket(engine._gflat, world, average, group)
        return engine._gflat
    cur = torch.cuda.current_stream()
    comm = comm_stream or _comm_stream(engine._gflat.device)
    buckets = engine.grad_buckets
    with torch.cuda.stream(comm):
        for l in reversed(range(len(buckets) - 1)):
            lo, hi = buckets[l]
            if hi <= lo:
                continue
            comm.wait_event(engine.layer_events[l])
            _reduce_bucket(engine._gflat[lo:hi], world, average, group)
        comm.wait_stream(cur)                               # shared modules: final when the backward's own stream is
        lo, hi = buckets[-1]
        if hi > lo:
            _reduce_bucket(engine._gflat[lo:hi], world, average, group)
    cur.wait_stream(comm)
    return engine._gflat


_COMM_STREAMS = {}


def _comm_stream(device):
    key = str(device)
    if key not in _COMM_STREAMS:
        _COMM_STREAMS[key] = torch.cuda.Stream(device=device)
    return _COMM_STREAMS[key]


def _reduce_bucket(t, world, average, group):
    if average and t.is_cuda and dist.get_backend(group) == "nccl":
        dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)       # NCCL averages inside the collective: no division pass
        return
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    if average:
        t.div_(world)
