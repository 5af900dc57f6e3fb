#!/usr/bin/env python
"""Benchmark of the detection hot path (BASELINE.json metric: images/sec at 608x608, T=10 MC samples).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 1..5] [--precision fp16|fp16x3|fp32]

One step = one pass of the hot path (backbone once + head x T + decode + NMS(1000)) over one batch of synthetic images
per GPU; weights are random-init (byolo.weights.synthetic, seed 0).  --config selects the workload, numbered as in
SURVEY.md 8d (= BASELINE.json configs[n-1]); the default 3 is the configuration the metric is quoted on:

  1  configs[0]  standard YOLOv3, 1 x 608x608, T=1 (the reference's own CPU-runnable case)
  2  configs[1]  aleatoric head, batch 8, 608x608
  3  configs[2]  epistemic MC-dropout T=10, batch 16 per GPU, 608x608                       (default, headline)
  4  configs[3]  epistemic MC-dropout T=30, 416x416, 4 images per GPU (batch 32 on 8 GPUs), gathered
  5  configs[4]  NMS stress: rows [8 per GPU, 22743, 23] with ties and planted clusters, cap 1000 (batch 64 on 8 GPUs)

Prints ONE JSON line (contract in the task statement): `value` = device-resident throughput of exactly K steps, timed
after >= 1.5 s of the same load so that the power-capped SM clock has settled (the sustained rate, not the boost);
`e2e` = the same metric through byolo_submit_host / byolo_wait_host (pinned host images in, host detections out,
copies inside the timed region); `roofline` for the dominant kernel (tcgen05 conv stack, tensor bound; config 5: the
NMS kernel, HBM bound); `cpu_baseline` = the oracle port on the host cores (rank 0).  With N > 1 ranks the images are
sharded and every step ends in the one all-gather of the packed detections (byolo.dist.ShardedDetector - the product
path, issued asynchronously into two buffer pairs).
`--impl reference` times the reference arm: the CPU restatement of the reference path (TensorFlow 1.x is not
installable here) on all host cores, model built once outside the timed loop.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time



def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# torchrun exports OMP_NUM_THREADS=1 to every rank; rank 0 also times the CPU baseline / reference arm on ALL host cores, so
# its OpenMP pool must be sized before numpy / torch load their runtimes (torch.set_num_threads alone does not resize oneDNN's)
if os.environ.get('RANK', '0') == '0':
    os.environ['OMP_NUM_THREADS'] = os.environ['MKL_NUM_THREADS'] = str(host_threads())

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'bayesian-yolov3_b200'), os.path.join(ROOT, 'tests')]

METRIC = 'images/sec at 608x608, T=10 MC samples (epistemic, incl. decode + NMS)'
CONFIGS = {
    1: dict(workload='configs[0]: standard YOLOv3, 1x608x608, T=1, cls_cnt 2, NMS cap 1000',
            variant='standard', img=608, T=1, batch_per_gpu=1, cls_cnt=2, max_out=1000),
    2: dict(workload='configs[1]: aleatoric head, batch 8/GPU, 608x608, cls_cnt 2, NMS cap 1000',
            variant='aleatoric', img=608, T=1, batch_per_gpu=8, cls_cnt=2, max_out=1000),
    3: dict(workload='configs[2]: epistemic MC-dropout T=10, batch 16/GPU, 608x608, cls_cnt 2, NMS cap 1000',
            variant='epistemic', img=608, T=10, batch_per_gpu=16, cls_cnt=2, max_out=1000),
    4: dict(workload='configs[3]: epistemic MC-dropout T=30, 4 images/GPU (batch 32 on 8 GPUs), 416x416, cls_cnt 2, NMS cap 1000, gathered',
            variant='epistemic', img=416, T=30, batch_per_gpu=4, cls_cnt=2, max_out=1000),
    5: dict(workload='configs[4]: NMS stress, rows [8/GPU (64 on 8 GPUs), 22743, 23], 5% score ties, 200 clusters of 30, cap 1000, gathered',
            variant='nms-stress', img=608, T=1, batch_per_gpu=8, cls_cnt=2, max_out=1000),
}
METRICS = {1: 'images/sec at 608x608, standard YOLOv3 (incl. decode + NMS)', 2: 'images/sec at 608x608, aleatoric head (incl. decode + NMS)',
           3: METRIC, 4: 'images/sec at 416x416, T=30 MC samples (epistemic, incl. decode + NMS + all-gather)',
           5: 'images/sec of class-agnostic NMS(1000) + gather on 22743 x 23 rows'}
OBJ = {'standard': 4, 'aleatoric': 9, 'epistemic': 14, 'nms-stress': 14}


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return float(p['bf16_tflops_sustained']), float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json: sustained bf16, copy bandwidth)'
    except Exception:
        return 1400.0, 6650.0, 'fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained, 6.65 TB/s)'


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


class CpuReference:
    """The oracle port of the reference path on the host cores: forward (torch CPU fp32) + decode + NMS, or the NMS alone
    for the stress configuration.  Model construction (weight synthesis, kernel layout) happens ONCE, here; run() is what
    gets timed.  torchrun exports OMP_NUM_THREADS=1: the thread count is set explicitly to all cores of the box."""

    def __init__(self, cfg, reserve=0):
        import torch
        from byolo import priors as P, weights as W
        from oracle import net as ON
        # reserve: cores left to the other ranks of a multi-GPU run, which spin in the closing barrier meanwhile
        self.cfg, self.threads = cfg, max(1, host_threads() - reserve)
        torch.set_num_threads(self.threads)
        self.pri = P.as_scale_list(P.by_stride('ECP_9_PRIORS'))
        v = cfg['variant']
        if v == 'nms-stress':
            from golden_inputs import stress_rows
            self.rows = stress_rows(4, 105)
        else:
            self.fwd = ON.Forward(v, W.synthetic(v, cfg['cls_cnt'], 0), cfg['cls_cnt'], torch.float32)
            self.imgs = np.random.default_rng(103).random((4, cfg['img'], cfg['img'], 3), dtype=np.float32)

    def run(self, n_images, first=0):
        """Processes n_images images (cycling over 4 distinct inputs); returns nothing - the caller times it."""
        from oracle import decode as D, nms as ONMS
        cfg, v = self.cfg, self.cfg['variant']
        for i in range(first, first + n_images):
            if v == 'nms-stress':
                ONMS.nms_gather(self.rows[i % 4], OBJ[v], cfg['max_out'])
                continue
            res = self.fwd.run(self.imgs[i % 4:i % 4 + 1], T=cfg['T'] if v == 'epistemic' else None, seed=1003, image_index0=i)
            rows = D.rows_from_raw(v, res[0]['raw'], self.pri)
            rows = rows if v == 'epistemic' else rows[0]
            ONMS.nms_gather(rows, OBJ[v], cfg['max_out'])

    def describe(self, n):
        what = 'NMS(1000)+gather (C restatement of the TF kernel)' if self.cfg['variant'] == 'nms-stress' else \
            'forward (torch CPU fp32) + numpy decode + C NMS'
        return '%d image(s) of the same workload per step through the oracle port: %s; model built once outside the timed loop' % (n, what)


def run_reference(args, cfg, emit):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    ref = CpuReference(cfg)
    per_step = 64 if cfg['variant'] == 'nms-stress' else 1          # bounded sample per step
    for i in range(max(args.warmup, 1)):
        ref.run(per_step, i * per_step)
    t0 = time.perf_counter()
    for i in range(args.steps):
        ref.run(per_step, i * per_step)
    dt = time.perf_counter() - t0
    v = per_step * args.steps / dt
    emit({'impl': 'reference', 'metric': METRICS[args.config], 'value': v, 'unit': 'images/s', 'n_gpus': args.gpus,
          'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
          'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
          'config': cfg,
          'cpu_baseline': {'value': v, 'unit': 'images/s', 'cores': ref.threads, 'kind': 'port', 'sample': ref.describe(per_step)},
          'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--config', type=int, default=3, choices=sorted(CONFIGS))
    ap.add_argument('--precision', default='fp16', choices=['fp16', 'fp16x3', 'fp32', 'fp16-simt'])
    ap.add_argument('--batch', type=int, default=0, help='images per GPU (default: the configuration\'s)')
    ap.add_argument('--warm-seconds', type=float, default=1.5, help='minimum duration of the untimed warm-up load')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-sustained', action='store_true', help='skip the extra 2 s sustained-rate pass')
    ap.add_argument('--no-parity-mode', action='store_true', help='skip the side measurement of the split-fp16 (fp16x3) mode')
    ap.add_argument('--layers', action='store_true', help='print the per-launch table to stderr')
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.batch:
        cfg['batch_per_gpu'] = args.batch
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
        return run_reference(args, cfg, emit)

    import torch
    import torch.distributed as dist
    import byolo
    from byolo import dist as bd, weights as W

    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    B, T, S, K, Wm, MO = cfg['batch_per_gpu'], cfg['T'], cfg['img'], args.steps, max(args.warmup, 3), cfg['max_out']
    variant = cfg['variant']
    stress = variant == 'nms-stress'
    rng = np.random.default_rng(1000 + rank)
    n_rot = 4                                           # rotating inputs > 126 MB L2 in the conv configurations
    if stress:
        from golden_inputs import stress_rows
        D, N = 23, 22743
        host = [torch.from_numpy(stress_rows(B, 105 + 10 * rank + i)).pin_memory() for i in range(n_rot)]
        eng = None
    else:
        eng = byolo.Engine(variant, (S, S), cfg['cls_cnt'], T=T, max_batch=B, precision=args.precision)
        eng.load_weights(W.synthetic(variant, cfg['cls_cnt'], 0))
        D, N = eng.D, eng.N
        host = [torch.from_numpy(rng.random((B, S, S, 3), dtype=np.float32)).pin_memory() for _ in range(n_rot)]
    devs = [h.to(dev) for h in host]
    in_bytes = host[0].numel() * 4
    l2_note = '%d rotating input batches (%.0f MB)%s' % (n_rot, n_rot * in_bytes / 1e6,
                                                        ' > L2; per-step activations (GBs) stream through HBM' if not stress else
                                                        '; %s 126 MB L2' % ('>' if n_rot * in_bytes > 126e6 else 'inputs fit the'))
    flush = None
    if stress and n_rot * in_bytes <= 140e6:            # small inputs: flush L2 between the timed iterations instead
        flush = torch.empty(160 * 1024 * 1024, dtype=torch.uint8, device=dev)
        l2_note += ', L2 flushed (160 MB memset) before every timed step'

    # ---- the step: local hot path (+ the one all-gather when sharded) ----
    if stress:
        def run_packed(rows, i0, out):
            _lib_nms(rows, out)

        def _lib_nms(rows, out):
            from byolo import _lib
            from byolo.engine import _ptr, _stream
            _lib.check(_lib.lib().byolo_nms_ex(_ptr(rows), rows.shape[0], N, D, OBJ[variant], 0.5, MO, _ptr(out), None, None, 1, 0, 0, _stream()))
    else:
        def run_packed(img, i0, out):
            eng.detect_packed(img, seed=1003, image_index0=i0, max_out=MO, out=out)
    sd = None
    packed = [torch.zeros((B, MO + 1, D), dtype=torch.float32, device=dev) for _ in range(2)]
    if world > 1:
        sd = bd.ShardedDetector(run_packed, world * B, D, max_out=MO, device=dev)

    def step(i):
        if flush is not None:
            flush.zero_()
        if sd is not None:                              # the product's sharded path: detect_packed -> async ncclAllGather
            sd.submit(devs[i % n_rot], i % 2)
        else:
            run_packed(devs[i % n_rot], rank * B, packed[i % 2])

    def drain():
        if sd is not None:
            sd.wait(0)
            sd.wait(1)

    def barrier():
        drain()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up: >= W steps AND >= warm_seconds of the same load (the power cap settles within ~0.1-1 s) ----
    # Every rank must issue the SAME number of steps (each step carries a collective): rank 0 times a probe and broadcasts
    # the count; nothing below depends on a per-rank clock.
    for i in range(Wm):
        step(i)
    barrier()
    t_w0 = time.perf_counter()
    for i in range(4):
        step(i)
    barrier()
    per_step_s = max((time.perf_counter() - t_w0) / 4, 1e-5)
    n_more = torch.tensor([int(min(max(args.warm_seconds, 0.0) / per_step_s, 20000))], dtype=torch.int64, device=dev)
    if world > 1:
        dist.broadcast(n_more, src=0)
    n_more = int(n_more[0])
    for i in range(n_more):
        step(i)
        if i % 8 == 7:
            torch.cuda.synchronize()
    n_warm = Wm + 4 + n_more
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    for i in range(4):                                  # keep the load on while the sampler starts
        step(i)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if eng is not None:
        eng.profile(2)      # coarse: events after the stem and before the decode launch only - the conv launches in between
    barrier()               # overlap as always, and their total duration is measured inside the timed region
    ev0.record()
    for i in range(K):
        step(i)
    drain()                 # the current stream waits for the outstanding gathers: ev1 sits behind them
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    coarse = eng.profile_read_coarse()[-K:] if eng is not None else np.zeros((0, 3))
    prof, ms_prof = [], None
    if eng is not None:
        eng.profile(False)
        # Second pass over the same K steps with the library's per-launch CUDA events switched on (per-layer table).  The
        # events sit between the launches, which also keeps a launch from overlapping the tail of its predecessor
        # (programmatic dependent launch), so this pass is a little slower than the headline one; both are reported.
        eng.profile(True)
        pv0, pv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        pv0.record()
        for i in range(K):
            step(i)
        drain()
        pv1.record()
        barrier()
        ms_prof = pv0.elapsed_time(pv1)
        prof = eng.profile_read()
        eng.profile(False)
    nms_ms = None
    if stress:              # per-launch duration of the NMS kernel alone (events around each launch, L2 flushed before)
        ts = []
        for i in range(min(K, 20)):
            if flush is not None:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            run_packed(devs[i % n_rot], rank * B, packed[i % 2])
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        nms_ms = statistics.mean(ts)

    # ---- end to end through the host-buffer entry points (H2D of every step's inputs, D2H of its detections) ----
    ms_e2e = ms_e2e_wall = ms_serial = None
    out_bytes = B * MO * D * 4 + B * 4
    if eng is not None:
        # pipelined form: two slots, so the copy of batch i+1 overlaps the compute of batch i; every step still moves its
        # own inputs in and detections out inside the timed region and the last wait_host is inside it too.
        outs = [(torch.empty((B, MO, D), dtype=torch.float32).pin_memory(), torch.empty((B,), dtype=torch.int32).pin_memory())
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
        out = (np.empty((B, MO, D), np.float32), np.empty((B,), np.int32))
        eng.detect_host(host[0], seed=1003, image_index0=rank * B, out=out)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(min(K, 5)):
            eng.detect_host(host[i % n_rot], seed=1003, image_index0=rank * B, out=out)
        s1.record()
        barrier()
        ms_serial = s0.elapsed_time(s1) / min(K, 5)
    else:
        # NMS stress end to end: host rows in (pinned), packed detections out, serial per step on the current stream
        out_bytes = B * (MO + 1) * D * 4
        hout = torch.empty((B, MO + 1, D), dtype=torch.float32).pin_memory()
        din = torch.empty_like(devs[0])
        barrier()
        t_host0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            din.copy_(host[i % n_rot], non_blocking=True)
            run_packed(din, rank * B, packed[0])
            hout.copy_(packed[0], non_blocking=True)
        e1.record()
        barrier()
        ms_e2e_wall = (time.perf_counter() - t_host0) * 1e3
        ms_e2e = e0.elapsed_time(e1)
    # sustained rate: the same step back to back for >= 2 s
    n_sus, t_sus = 0, 0.0
    if not args.no_sustained:
        ms_max = torch.tensor([ms / K], device=dev)
        if world > 1:
            dist.all_reduce(ms_max, op=dist.ReduceOp.MAX)       # every rank must run the same number of steps (all-gather inside)
        n_sus = max(10, int(2000.0 / max(float(ms_max[0]), 1e-3)))
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        u0.record()
        for i in range(n_sus):
            step(i)
        drain()
        u1.record()
        barrier()
        t_sus = u0.elapsed_time(u1)
    clocks = sampler.stop()
    # ---- the same K steps in the split-fp16 mode (the mode that meets the 1e-3 element-wise parity bound), side by side ----
    x3 = None
    if eng is not None and args.precision == 'fp16' and not args.no_parity_mode and world == 1:
        eng3 = byolo.Engine(variant, (S, S), cfg['cls_cnt'], T=T, max_batch=B, precision='fp16x3')
        eng3.load_weights(W.synthetic(variant, cfg['cls_cnt'], 0))
        t_w0, n3 = time.perf_counter(), 0
        while n3 < Wm or (time.perf_counter() - t_w0) < args.warm_seconds:
            eng3.detect_packed(devs[n3 % n_rot], seed=1003, image_index0=rank * B, max_out=MO, out=packed[n3 % 2])
            n3 += 1
            if n3 % 4 == 0:
                torch.cuda.synchronize()
        eng3.profile(2)
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        x0.record()
        for i in range(K):
            eng3.detect_packed(devs[i % n_rot], seed=1003, image_index0=rank * B, max_out=MO, out=packed[i % 2])
        x1.record()
        torch.cuda.synchronize()
        ms3 = x0.elapsed_time(x1)
        c3 = eng3.profile_read_coarse()[-K:]
        eng3.profile(False)
        # end to end through the same host-buffer entry points
        for i in range(2):
            eng3.submit_host(host[i % n_rot], outs[i % 2], i % 2, seed=1003, image_index0=rank * B)
        eng3.wait_host(0)
        eng3.wait_host(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(K):
            if i >= 2:
                eng3.wait_host(i % 2)
            eng3.submit_host(host[i % n_rot], outs[i % 2], i % 2, seed=1003, image_index0=rank * B)
        for i in range(max(K - 2, 0), K):
            eng3.wait_host(i % 2)
        torch.cuda.synchronize()
        ms3_e2e = (time.perf_counter() - t0) * 1e3
        stack3 = float(c3[:, 1].mean()) if len(c3) else None
        x3 = {'precision': 'fp16x3', 'what': 'split-fp16 tensor-core mode: every operand a hi+lo fp16 pair, 3 MMAs per K step, chunked fp32 accumulation; '
                                             'meets the 1e-3 element-wise parity bound (tests/test_gpu_fullsize.py, profiles/r02/parity_*_fp16x3.txt)',
              'value': B * K / (ms3 * 1e-3), 'unit': 'images/s', 'ms_per_step': ms3 / K, 'steps': K,
              'e2e': {'value': B * K / (ms3_e2e * 1e-3), 'unit': 'images/s'},
              'conv_stack_ms_per_step': stack3, 'gpu_launches': eng3.launch_count(B) * K,
              'mma_factor': eng3.flops_per_image(executed=True) / eng3.flops_per_image()}
        eng3.close()
    if world > 1:
        t = torch.tensor([ms, ms_e2e, t_sus, ms_e2e_wall], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, t_sus, ms_e2e_wall = (float(x) for x in t)

    if rank == 0:
        peak_tf, peak_gbs, peak_src = peaks()
        dtype = {'fp32': 'f32', 'fp16': 'f16', 'fp16-simt': 'f16', 'fp16x3': 'f16x3 (split fp16: hi+lo operand pairs, fp32 accumulate)'}[args.precision]
        res = {'metric': METRICS[args.config], 'value': world * B * K / (ms * 1e-3), 'unit': 'images/s', 'n_gpus': world, 'steps': K,
               'warmup': Wm, 'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
               'dtype': 'f32 (IoU compares)' if stress else dtype, 'data': 'synthetic', 'config': cfg,
               'precision': None if stress else args.precision, 'l2': l2_note,
               'warmup_steps_run': n_warm, 'warmup_seconds': args.warm_seconds,
               'timing_note': 'K steps timed with CUDA events after >= %.1f s of the same load: the power cap (sw_power_cap, ~1000 W) has pulled the SM '
                              'clock to its sustained level by then; round-1 lines were timed 5 steps after a cold start (boost clock) and read '
                              '~7 %% higher for the same kernels' % args.warm_seconds,
               'p50_ms_per_img': ms / K / B, 'ms_per_step_with_per_launch_events': (ms_prof / K) if ms_prof else None,
               'sustained': ({'value': world * B * n_sus / (t_sus * 1e-3), 'unit': 'images/s', 'steps': n_sus, 'seconds': t_sus * 1e-3}
                             if n_sus else None),
               'clocks': clocks,
               'e2e': {'value': world * B * K / (max(ms_e2e, ms_e2e_wall) * 1e-3), 'unit': 'images/s',
                       'h2d_bytes_per_step': in_bytes, 'd2h_bytes_per_step': out_bytes,
                       'api': 'byolo_submit_host/byolo_wait_host, 2 slots, pinned host buffers' if eng is not None else
                              'pinned host rows -> cudaMemcpyAsync -> byolo_nms_ex(packed) -> cudaMemcpyAsync, serial',
                       'serial_detect_host_ms_per_step': ms_serial},
               'gpu_launches': (eng.launch_count(B) if eng is not None else 1) * K,
               'multi_gpu': ({'gather': 'byolo.dist.ShardedDetector: byolo_detect_packed -> one async ncclAllGather of [%d,%d,%d] fp32 per rank and step '
                                        '(detections + count row), two buffer pairs' % (B, MO + 1, D)} if world > 1 else None)}
        if eng is not None:
            conv = [p for p in prof if p['kind'] == 'conv']
            conv_ms = sum(p['ms'] for p in conv)
            conv_fl = sum(p['flops'] for p in conv)
            step_ms = sum(p['ms'] for p in prof)
            achieved_per_launch = conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
            stack_ms = float(coarse[:, 1].mean()) if len(coarse) else 0.0      # conv stack of the timed steps (CUDA events, launch stream)
            achieved = conv_fl / (stack_ms * 1e-3) / 1e12 if stack_ms > 0 else achieved_per_launch
            traffic = None                              # DRAM bytes of the conv launches of one step, from the committed ncu capture
            for name in ('profiles/r02/ncu_traffic.json', 'profiles/r01/ncu_v8_traffic.json'):
                try:
                    with open(os.path.join(ROOT, name)) as f:
                        tj = json.load(f)
                    if args.config == 3 and args.precision == tj.get('precision', 'fp16') and B == 16 and tj['conv_launches'] == len(conv):
                        traffic = tj['conv_dram_bytes_per_step']
                        break
                except Exception:
                    pass
            mma_factor = eng.flops_per_image(executed=True) / eng.flops_per_image()
            res['roofline'] = {
                'bound': 'tensor', 'kernel': 'conv_umma_kernel (all %d conv launches of a step)' % len(conv),
                'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved / peak_tf,
                'peak_source': peak_src, 'traffic': traffic,
                'how': 'sum of the ALGORITHMIC FLOPs (2*MAC of the reference graph) of the conv launches / their duration in the timed steps '
                       '(events after the stem and before the decode launch, mean over the steps)',
                'executed_over_algorithmic_flops': mma_factor,
                'tensor_pipe_frac_executed': achieved * mma_factor / peak_tf,
                'conv_stack_ms_per_step': stack_ms, 'stem_ms_per_step': float(coarse[:, 0].mean()) if len(coarse) else None,
                'decode_nms_ms_per_step': float(coarse[:, 2].mean()) if len(coarse) else None,
                'achieved_with_per_launch_events': achieved_per_launch,
                'conv_share_of_step': (stack_ms / (ms / K)) if stack_ms else (conv_ms / step_ms if step_ms else None),
                'flops_per_image': eng.flops_per_image(), 'step_tflops': eng.flops_per_image() * B / (ms / K * 1e-3) / 1e12}
            res['breakdown_ms'] = {k: sum(p['ms'] for p in prof if p['kind'] == k) for k in ('stem', 'conv', 'decode', 'nms')}
            if x3 is not None:
                if x3['conv_stack_ms_per_step']:
                    alg = conv_fl / (x3['conv_stack_ms_per_step'] * 1e-3) / 1e12
                    x3['roofline'] = {'bound': 'tensor', 'achieved': alg, 'executed': x3['mma_factor'] * alg, 'peak': peak_tf, 'unit': 'TFLOP/s',
                                      'frac': alg / peak_tf, 'tensor_pipe_frac_executed': x3['mma_factor'] * alg / peak_tf}
                res['parity_mode'] = x3
            big = [p for p in conv if p['ms'] > 0.3 and p['sm_mhz'] > 0]
            if big:      # effective SM clock inside the long conv launches (clock64/globaltimer): shows power-cap throttling
                res['clocks']['sm_mhz_in_conv_kernels'] = sum(p['sm_mhz'] * p['ms'] for p in big) / sum(p['ms'] for p in big)
            if args.layers:
                for p in prof:
                    tf = p['flops'] / (p['ms'] * 1e-3) / 1e12 if p['ms'] > 0 else 0
                    print('%-6s layer %3d  %8.3f ms  %8.1f TFLOP/s  %6.0f MHz' % (p['kind'], p['layer'], p['ms'], tf, p['sm_mhz']), file=sys.stderr)
        else:
            # NMS: algorithmic bytes per image = 22743 x (4 + 1) x 4 B read + 1000 x 23 x 4 B written
            alg = B * (N * 5 * 4 + MO * D * 4)
            ach = alg / (nms_ms * 1e-3) / 1e9
            res['roofline'] = {'bound': 'hbm', 'kernel': 'nms_kernel (one launch per step, %d images)' % B, 'achieved': ach, 'peak': peak_gbs,
                               'unit': 'GB/s', 'frac': ach / peak_gbs, 'peak_source': peak_src, 'traffic': None,
                               'how': 'algorithmic bytes (N x 5 x 4 B read + 1000 x D x 4 B written per image) / mean launch duration (CUDA events '
                                      'around each launch, L2 flushed before); the kernel is latency bound (sort + sequential greedy scan), not byte bound',
                               'nms_ms_per_launch': nms_ms}
        if not args.no_cpu_baseline:
            ref = CpuReference(cfg, reserve=world - 1)
            n_cpu = 256 if stress else (8 if T >= 10 else 16)      # ~5-20 s of CPU work
            if world > 1 and not stress:
                n_cpu = 4                                           # the other ranks wait (busy) in the closing barrier
            ref.run(1)
            t0 = time.perf_counter()
            ref.run(n_cpu)
            v = n_cpu / (time.perf_counter() - t0)
            res['cpu_baseline'] = {'value': v, 'unit': 'images/s', 'cores': ref.threads, 'kind': 'port', 'sample': ref.describe(n_cpu)}
        emit(res)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
