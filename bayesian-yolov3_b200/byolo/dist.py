"""Multi-GPU: images shard over ranks (one process per GPU, weights replicated), every image is independent
(SURVEY.md 8e); the only exchange is ONE all-gather of the final padded detections + counts.  Dropout masks are keyed
by the GLOBAL image index, so the gathered result is identical to a single-GPU run over the same images.
Host-side logic only; the transport is torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_images, rank, world):
    """Contiguous block of images for `rank`; earlier ranks take the remainder."""
    base, rem = divmod(n_images, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_detections(boxes, counts, n_images, group=None):
    """boxes [b_local,max_out,D] fp32, counts [b_local] int32 on this rank -> ([n_images,max_out,D], [n_images]) on every
    rank, in global image order.  Ranks may hold different b_local (uneven shards are padded to the largest shard so
    that one fixed-size all-gather suffices; counts ride in the same message as one extra row per image)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    max_out, D = boxes.shape[1], boxes.shape[2]
    per = -(-n_images // world)                                       # ceil: largest shard
    msg = torch.zeros((per, max_out + 1, D), dtype=torch.float32, device=boxes.device)
    b_local = boxes.shape[0]
    assert (shard_range(n_images, rank, world)[1] - shard_range(n_images, rank, world)[0]) == b_local
    msg[:b_local, :max_out] = boxes
    msg[:b_local, max_out, 0] = counts.to(torch.float32)              # exact for counts < 2^24
    out = torch.empty((world,) + tuple(msg.shape), dtype=torch.float32, device=boxes.device)
    if dist.get_backend(group) == 'nccl':
        dist.all_gather_into_tensor(out, msg, group=group)               # one ncclAllGather
    else:
        dist.all_gather(list(out.unbind(0)), msg, group=group)           # gloo (CPU tests)
    parts_b, parts_c = [], []
    for r in range(world):
        s, e = shard_range(n_images, r, world)
        parts_b.append(out[r, :e - s, :max_out])
        parts_c.append(out[r, :e - s, max_out, 0].to(torch.int32))
    return torch.cat(parts_b), torch.cat(parts_c)


class ShardedDetector:
    """detect(images) over all ranks: `run_local(images_local, image_index0)` -> (boxes, counts) is the per-rank hot
    path (Engine.detect on a GPU; a stub in the CPU tests)."""

    def __init__(self, run_local, group=None):
        self.run_local, self.group = run_local, group

    def detect(self, images):
        """images: the GLOBAL batch [n,H,W,3] (every rank passes the same tensor or at least its own slice filled)."""
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        n = images.shape[0]
        s, e = shard_range(n, rank, world)
        boxes, counts = self.run_local(images[s:e], s)
        return gather_detections(boxes, counts, n, self.group)
