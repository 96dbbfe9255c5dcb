"""Multi-GPU plumbing: one process per GPU, batches sharded by image, no data-path collective except the final
gather of the fixed-size detection records.  Works with the NCCL backend on GPUs and with gloo
on CPU tensors (used by the world_size-2 CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(global_batch, rank, world_size):
    """Contiguous split of the global batch by image index: rank r owns images [lo, hi)."""
    if global_batch % world_size != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world_size}")
    per = global_batch // world_size
    return rank * per, (rank + 1) * per


def pack_detections(boxes, classes, scores, num_valid):
    """[B,max,4] f32, [B,max] i64, [B,max] f32, [B] i32 -> one float32 record tensor [B, max*6 + 1]
    (x1,y1,x2,y2,score,class per slot, then num_valid).  Classes < 2^24 are exact in float32."""
    B, mx = scores.shape
    rec = torch.empty((B, mx * 6 + 1), dtype=torch.float32, device=scores.device)
    r = rec[:, :mx * 6].view(B, mx, 6)
    r[..., 0:4] = boxes
    r[..., 4] = scores
    r[..., 5] = classes.to(torch.float32)
    rec[:, mx * 6] = num_valid.to(torch.float32)
    return rec


def unpack_detections(rec):
    B, w = rec.shape
    mx = (w - 1) // 6
    r = rec[:, :mx * 6].view(B, mx, 6)
    return r[..., 0:4], r[..., 5].to(torch.int64), r[..., 4], rec[:, mx * 6].to(torch.int32)


def gather_packed(rec, group=None):
    """All-gather of already packed detection records ``[B, max*6 + 1]`` (``Detector.detections_graphed(x, packed=True)``):
    one collective, no other launches; ``unpack_detections`` of the result is views plus two small casts.
    Single-process: returns ``rec``."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return rec
    world = dist.get_world_size(group)
    out = torch.empty((world * rec.shape[0], rec.shape[1]), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, rec.contiguous(), group=group)
    return out


def gather_detections(boxes, classes, scores, num_valid, group=None):
    """All ranks end up with the detections of the whole global batch, ordered by global image index (rank-major,
    which is image order for ``shard_range`` shards).  Single-process: returns the inputs."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return boxes, classes, scores, num_valid
    rec = pack_detections(boxes, classes, scores, num_valid).contiguous()
    world = dist.get_world_size(group)
    out = torch.empty((world * rec.shape[0], rec.shape[1]), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, rec, group=group)
    return unpack_detections(out)
