#!/usr/bin/env python
"""Benchmark of the detection hot path (BASELINE.json metric: images/sec at 608x608, T=10 MC samples).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One step = one pass of the hot path (backbone once + head x T + decode + NMS(1000)) over one batch of B=16 synthetic
608x608 images per GPU (BASELINE.json configs[2]); weights are random-init (byolo.weights.synthetic, seed 0).
Prints ONE JSON line (contract in the task statement): `value` = device-resident throughput, `e2e` = the same metric
through byolo_detect_host (pinned host images in, host detections out, copies inside the timed region),
`roofline` for the dominant kernel (tcgen05 conv stack, tensor bound; per-launch CUDA events of a second pass over the
same steps), `cpu_baseline` = the oracle port on host cores.
`--impl reference` times the reference arm: the CPU restatement of the reference path (TensorFlow 1.x is not
installable here) on all host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'bayesian-yolov3_b200')]

METRIC = 'images/sec at 608x608, T=10 MC samples (epistemic, incl. decode + NMS)'
WORKLOAD = dict(workload='configs[2]: epistemic MC-dropout T=10, batch 16/GPU, 608x608, cls_cnt 2, NMS cap 1000',
                variant='epistemic', img=608, T=10, batch_per_gpu=16, cls_cnt=2, max_out=1000)


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return float(p['bf16_tflops_sustained']), float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json, sustained bf16)'
    except Exception:
        return 1400.0, 6650.0, 'fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(',')])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith('active') for r in self.rows)]
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


def cpu_oracle_rate(n_images, T, img, threads=None):
    """Oracle port of the reference path on host cores: forward (torch CPU fp32) + decode + NMS.  Returns img/s."""
    import torch
    from byolo import priors as P, weights as W
    from oracle import decode as D, net as ON, nms as ONMS
    if threads:
        torch.set_num_threads(threads)
    pri = P.as_scale_list(P.by_stride('ECP_9_PRIORS'))
    w = W.synthetic('epistemic', 2, 0)
    fwd = ON.Forward('epistemic', w, 2, torch.float32)
    imgs = np.random.default_rng(103).random((n_images, img, img, 3), dtype=np.float32)
    t0 = time.perf_counter()
    for b in range(n_images):
        res = fwd.run(imgs[b:b + 1], T=T, seed=1003, image_index0=b)
        rows = D.rows_from_raw('epistemic', res[0]['raw'], pri)
        ONMS.nms_gather(rows, 14, 1000)
    dt = time.perf_counter() - t0
    return n_images / dt, torch.get_num_threads()


def run_reference(args, emit):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    per_step = 1                                       # bounded sample: 1 image (T=10) per step, ~0.6 s on 8 cores
    for _ in range(max(args.warmup, 1)):
        cpu_oracle_rate(per_step, WORKLOAD['T'], WORKLOAD['img'])
    t0 = time.perf_counter()
    cores = 0
    for _ in range(args.steps):
        _, cores = cpu_oracle_rate(per_step, WORKLOAD['T'], WORKLOAD['img'])
    dt = time.perf_counter() - t0
    v = per_step * args.steps / dt
    sample = '%d image(s)/step of the same workload (608x608, T=10, decode+NMS), oracle port on %d host threads' % (per_step, cores)
    emit({'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'images/s', 'n_gpus': args.gpus,
                      'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
                      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                      'config': WORKLOAD,
                      'cpu_baseline': {'value': v, 'unit': 'images/s', 'cores': cores, 'kind': 'port', 'sample': sample},
                      'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--precision', default='fp16')
    ap.add_argument('--batch', type=int, default=WORKLOAD['batch_per_gpu'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-sustained', action='store_true', help='skip the 2 s sustained-rate pass')
    ap.add_argument('--layers', action='store_true', help='print the per-launch table to stderr')
    args = ap.parse_args()
    # stdout carries exactly ONE line (the JSON record): whatever libraries print (e.g. "NCCL version ...") goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(record):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(record), flush=True)
        os.dup2(2, 1)

    if args.impl == 'reference':
        return run_reference(args, emit)

    import torch
    import torch.distributed as dist
    import byolo
    from byolo import weights as W

    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    B, T, S, K, Wm = args.batch, WORKLOAD['T'], WORKLOAD['img'], args.steps, max(args.warmup, 3)
    eng = byolo.Engine('epistemic', (S, S), 2, T=T, max_batch=B, precision=args.precision)
    eng.load_weights(W.synthetic('epistemic', 2, 0))
    rng = np.random.default_rng(1000 + rank)
    n_rot = 4                                           # 4 x 71 MB of images > 126 MB L2: inputs never L2 resident
    host = [torch.from_numpy(rng.random((B, S, S, 3), dtype=np.float32)).pin_memory() for _ in range(n_rot)]
    devs = [h.to(dev) for h in host]
    gathered = torch.empty((world, B, 1000, eng.D), dtype=torch.float32, device=dev) if world > 1 else None

    def step(i):
        boxes, cnt, idx = eng.detect(devs[i % n_rot], seed=1003, image_index0=rank * B)
        if world > 1:                                   # the one exchange of the path: final detections
            dist.all_gather_into_tensor(gathered, boxes)
        return boxes, cnt

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(Wm):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.profile(2)          # coarse: events after the stem and before the decode launch only - the 74 conv launches in between
    barrier()               # overlap as always, and their total duration is measured inside the timed region
    ev0.record()
    for i in range(K):
        boxes, cnt = step(i)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    coarse = eng.profile_read_coarse()[-K:]
    eng.profile(False)
    # Second pass over the same K steps with the library's per-launch CUDA events switched on (roofline inputs).  The
    # events sit between the launches, which also keeps a launch from overlapping the tail of its predecessor
    # (programmatic dependent launch), so this pass is a little slower than the headline one; both are reported.
    eng.profile(True)
    pv0, pv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    pv0.record()
    for i in range(K):
        step(i)
    pv1.record()
    barrier()
    ms_prof = pv0.elapsed_time(pv1)
    prof = eng.profile_read()
    eng.profile(False)

    # ---- end to end through the host-buffer entry points (H2D of every step's images, D2H of its detections) ----
    # pipelined form: two slots, so the copy of batch i+1 overlaps the compute of batch i; every step still moves its own
    # 71 MB in and 1.5 MB out inside the timed region and the last wait_host is inside it too.
    outs = [(torch.empty((B, 1000, eng.D), dtype=torch.float32).pin_memory(), torch.empty((B,), dtype=torch.int32).pin_memory())
            for _ in range(2)]
    for i in range(2):
        eng.submit_host(host[i % n_rot], outs[i % 2], i % 2, seed=1003, image_index0=rank * B)
    eng.wait_host(0)
    eng.wait_host(1)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    e0.record()
    for i in range(K):
        if i >= 2:
            eng.wait_host(i % 2)
        eng.submit_host(host[i % n_rot], outs[i % 2], i % 2, seed=1003, image_index0=rank * B)
    for i in range(max(K - 2, 0), K):
        eng.wait_host(i % 2)
    e1.record()
    barrier()
    ms_e2e_wall = (time.perf_counter() - t_host0) * 1e3     # barrier() above synchronised: wall time covers the last D2H
    ms_e2e = e0.elapsed_time(e1)
    # serial form (one call = copy in, compute, copy out, sync), for reference
    out = (np.empty((B, 1000, eng.D), np.float32), np.empty((B,), np.int32))
    eng.detect_host(host[0], seed=1003, image_index0=rank * B, out=out)
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(min(K, 5)):
        eng.detect_host(host[i % n_rot], seed=1003, image_index0=rank * B, out=out)
    s1.record()
    barrier()
    ms_serial = s0.elapsed_time(s1) / min(K, 5)
    # sustained rate: the same step back to back for >= 2 s (the power cap pulls the SM clock down within ~0.1 s of load;
    # K = 20 steps end before that has settled)
    n_sus, t_sus = 0, 0.0
    if not args.no_sustained:
        ms_max = torch.tensor([ms / K], device=dev)
        if world > 1:
            dist.all_reduce(ms_max, op=dist.ReduceOp.MAX)       # every rank must run the same number of steps (all-gather inside)
        n_sus = max(10, int(2000.0 / float(ms_max[0])))
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        u0.record()
        for i in range(n_sus):
            step(i)
        u1.record()
        barrier()
        t_sus = u0.elapsed_time(u1)
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms, ms_e2e, t_sus], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, t_sus = float(t[0]), float(t[1]), float(t[2])

    if rank == 0:
        peak_tf, peak_gbs, peak_src = peaks()
        conv = [p for p in prof if p['kind'] == 'conv']
        conv_ms = sum(p['ms'] for p in conv)
        conv_fl = sum(p['flops'] for p in conv)
        step_ms = sum(p['ms'] for p in prof)
        achieved_per_launch = conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        stack_ms = float(coarse[:, 1].mean()) if len(coarse) else 0.0      # conv stack of the timed steps (CUDA events, launch stream)
        achieved = conv_fl / (stack_ms * 1e-3) / 1e12 if stack_ms > 0 else achieved_per_launch
        traffic = None                                  # DRAM bytes of the conv launches of one step, from the committed ncu capture
        try:
            with open(os.path.join(ROOT, 'profiles', 'r01', 'ncu_v8_traffic.json')) as f:
                tj = json.load(f)
            if B == WORKLOAD['batch_per_gpu'] and tj['conv_launches'] == len(conv):
                traffic = tj['conv_dram_bytes_per_step']
        except Exception:
            pass
        res = {'metric': METRIC, 'value': world * B * K / (ms * 1e-3), 'unit': 'images/s', 'n_gpus': world, 'steps': K,
               'warmup': Wm, 'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
               'dtype': 'f16' if args.precision != 'fp32' else 'f32', 'data': 'synthetic',
               'config': dict(WORKLOAD, batch_per_gpu=B, precision=args.precision,
                              l2='4 rotating input batches (284 MB) > L2; per-step activations (GBs) stream through HBM'),
               'p50_ms_per_img': ms / K / B, 'ms_per_step_with_per_launch_events': ms_prof / K,
               'sustained': ({'value': world * B * n_sus / (t_sus * 1e-3), 'unit': 'images/s', 'steps': n_sus, 'seconds': t_sus * 1e-3}
                             if n_sus else None),
               'clocks': clocks,
               'e2e': {'value': world * B * K / (max(ms_e2e, ms_e2e_wall) * 1e-3), 'unit': 'images/s',
                       'h2d_bytes_per_step': B * S * S * 3 * 4, 'd2h_bytes_per_step': B * 1000 * eng.D * 4 + B * 4,
                       'api': 'byolo_submit_host/byolo_wait_host, 2 slots, pinned host buffers',
                       'serial_detect_host_ms_per_step': ms_serial},
               'gpu_launches': eng.launch_count(B) * K,
               'roofline': {'bound': 'tensor', 'kernel': 'conv_umma_kernel (all %d conv launches of a step)' % len(conv),
                            'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved / peak_tf,
                            'peak_source': peak_src, 'traffic': traffic,
                            'how': 'sum of the algorithmic FLOPs of the 74 conv launches / their duration in the timed steps (events after the stem '
                                   'and before the decode launch, mean over the steps)',
                            'conv_stack_ms_per_step': stack_ms, 'stem_ms_per_step': float(coarse[:, 0].mean()) if len(coarse) else None,
                            'decode_nms_ms_per_step': float(coarse[:, 2].mean()) if len(coarse) else None,
                            'achieved_with_per_launch_events': achieved_per_launch,
                            'conv_share_of_step': (stack_ms / (ms / K)) if stack_ms else (conv_ms / step_ms if step_ms else None),
                            'flops_per_image': eng.flops_per_image(), 'step_tflops': eng.flops_per_image() * B / (ms / K * 1e-3) / 1e12},
               'breakdown_ms': {k: sum(p['ms'] for p in prof if p['kind'] == k) for k in ('stem', 'conv', 'decode', 'nms')}}
        big = [p for p in conv if p['ms'] > 0.3 and p['sm_mhz'] > 0]
        if big:      # effective SM clock inside the long conv launches (clock64/globaltimer): shows power-cap throttling
            res['clocks']['sm_mhz_in_conv_kernels'] = sum(p['sm_mhz'] * p['ms'] for p in big) / sum(p['ms'] for p in big)
        if args.layers:
            for p in prof:
                tf = p['flops'] / (p['ms'] * 1e-3) / 1e12 if p['ms'] > 0 else 0
                print('%-6s layer %3d  %8.3f ms  %8.1f TFLOP/s  %6.0f MHz' % (p['kind'], p['layer'], p['ms'], tf, p['sm_mhz']), file=sys.stderr)
        if world == 1 and not args.no_cpu_baseline:
            n_cpu = 8                                   # ~5 s of CPU work on 8 cores
            v, cores = cpu_oracle_rate(n_cpu, T, S)
            res['cpu_baseline'] = {'value': v, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                                   'sample': '%d images of the same workload (608x608, T=10, decode+NMS) through the oracle '
                                             'port (torch CPU fp32 + numpy decode + C NMS)' % n_cpu}
        emit(res)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
