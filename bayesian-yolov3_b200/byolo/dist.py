"""Multi-GPU: images shard over ranks (one process per GPU, weights replicated), every image is independent
(SURVEY.md 8e); the only exchange is ONE all-gather of the final padded detections + counts.  Dropout masks are keyed
by the GLOBAL image index, so the gathered result is identical to a single-GPU run over the same images.

The message of a rank is the "packed" block the NMS kernel itself writes (byolo_detect_packed): [per, max_out + 1, D]
fp32, rows in selection order, zero padded, and (count, 0, ...) in row max_out of every image.  It goes from the kernel's
output buffer straight into ncclAllGather - no packing kernels, no torch ops; the gathered tensor is handed out as views.
The collective is issued asynchronously (its own NCCL stream) into one of `depth` buffer pairs, so the next batch's
kernels never wait for the slowest rank's previous batch.

Transport: torch.distributed (NCCL on GPUs; gloo in the CPU tests, where a stub fills the packed block)."""
import torch
import torch.distributed as dist


def shard_range(n_images, rank, world):
    """Contiguous block of images for `rank`; earlier ranks take the remainder."""
    base, rem = divmod(n_images, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class Gathered:
    """Result of one gather: `boxes` [n_images, max_out, D] and `counts` [n_images] (fp32, exact integers), both views of
    the receive buffer in global image order (a copy only when the shards are uneven)."""

    def __init__(self, recv, n_images, world, max_out):
        per = recv.shape[1]
        if n_images == per * world:                                     # even shards: pure views
            flat = recv.view(world * per, max_out + 1, recv.shape[3])
        else:                                                            # uneven: drop the padding images of the short ranks
            flat = torch.cat([recv[r, :shard_range(n_images, r, world)[1] - shard_range(n_images, r, world)[0]]
                              for r in range(world)])
        self.packed = flat
        self.boxes = flat[:, :max_out]
        self.counts = flat[:, max_out, 0]

    def counts_int(self):
        return self.counts.to(torch.int32)


class ShardedDetector:
    """detect(images) over all ranks.

    run_packed(images_local, image_index0, out) must fill `out` [b_local, max_out + 1, D] with the packed detections of
    this rank's images: `lambda img, i0, out: engine.detect_packed(img, seed, i0, out=out)` on a GPU, a stub in the CPU
    tests.  n_images is the fixed global batch size (buffers are allocated once)."""

    def __init__(self, run_packed, n_images, D, max_out=1000, device=None, group=None, depth=2):
        self.run_packed, self.group, self.n, self.D, self.max_out = run_packed, group, n_images, D, max_out
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.start, self.end = shard_range(n_images, self.rank, self.world)
        self.per = -(-n_images // self.world)                            # ceil: the largest shard sets the message size
        self.nccl = dist.get_backend(group) == 'nccl'
        # the tail images of a short rank stay zero for ever (count 0)
        self.send = [torch.zeros((self.per, max_out + 1, D), dtype=torch.float32, device=device) for _ in range(depth)]
        self.recv = [torch.empty((self.world, self.per, max_out + 1, D), dtype=torch.float32, device=device) for _ in range(depth)]
        self.work = [None] * depth

    def submit(self, images, slot=0):
        """images: the GLOBAL batch [n,H,W,3] (this rank reads its own slice) or already the local slice.  Enqueues the
        local detect and the all-gather of slot `slot`; returns without waiting for either."""
        self.wait(slot)                                                  # the previous gather into this slot has drained
        local = images[self.start:self.end] if images.shape[0] == self.n and self.world > 1 else images
        assert local.shape[0] == self.end - self.start
        b = local.shape[0]
        if b:
            self.run_packed(local, self.start, self.send[slot][:b])
        if self.nccl:
            self.work[slot] = dist.all_gather_into_tensor(self.recv[slot], self.send[slot], group=self.group, async_op=True)
        else:
            self.work[slot] = dist.all_gather(list(self.recv[slot].unbind(0)), self.send[slot], group=self.group, async_op=True)
        return slot

    def wait(self, slot=0):
        """Makes the current stream (NCCL) / the host (gloo) wait for the gather of `slot`."""
        if self.work[slot] is not None:
            self.work[slot].wait()
            self.work[slot] = None

    def result(self, slot=0):
        self.wait(slot)
        return Gathered(self.recv[slot], self.n, self.world, self.max_out)

    def detect(self, images):
        return self.result(self.submit(images, 0))
