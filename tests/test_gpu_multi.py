"""Multi-GPU parity: the gathered result of the image-sharded path == the single-GPU result, bit for
bit, through the product's own ShardedDetector (byolo_detect_packed -> one ncclAllGather).  Needs >= 2 GPUs
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`); skipped on a single-GPU box."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_images, size, T, ret):
    sys.path[:0] = [ROOT, os.path.join(ROOT, 'bayesian-yolov3_b200')]
    import torch.distributed as dist
    import byolo
    from byolo import dist as bd, weights as W
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world, device_id=dev)
    eng = byolo.Engine('epistemic', (size, size), 2, T=T, max_batch=n_images, precision='fp16').load_weights(W.synthetic('epistemic', 2, 0))
    imgs = torch.from_numpy(np.random.default_rng(7).random((n_images, size, size, 3), dtype=np.float32)).to(dev)
    sd = bd.ShardedDetector(lambda im, i0, out: eng.detect_packed(im.contiguous(), seed=5, image_index0=i0, out=out), n_images, eng.D,
                            device=dev)
    ok = True
    for slot in (0, 1, 0):
        got = sd.result(sd.submit(imgs, slot))
        torch.cuda.synchronize()
        single = eng.detect_packed(imgs, seed=5, image_index0=0)      # the same images on ONE GPU
        boxes, cnt, idx = eng.detect(imgs, seed=5, image_index0=0)
        torch.cuda.synchronize()
        ok = ok and torch.equal(torch.nan_to_num(got.packed), torch.nan_to_num(single))
        ok = ok and torch.equal(torch.nan_to_num(got.boxes), torch.nan_to_num(boxes)) and torch.equal(got.counts_int(), cnt)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize('n_images', [4, 5])
def test_gathered_equals_single_gpu_bit_for_bit(n_images):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    ret = mp.Manager().dict()
    port = 29600 + os.getpid() % 2000 + n_images
    mp.spawn(_worker, args=(2, port, n_images, 416, 3, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
