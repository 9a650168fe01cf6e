#!/usr/bin/env python
"""bench.py - headline benchmark of the ELD synthetic-noise training path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload noise|train] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic frames (4x512x512 packed raw,
SonyA7S2 calibrated parameters, fixed Philox seed).  Prints ONE JSON line (rank 0).
See DESIGN.md "measurement" for the definitions of value / e2e / roofline / cpu_baseline.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

SONY = (2.2881136684755243, 6.4508722699636545, 15583, 208.9766365993794)
FRAME_BYTES = 4 * 512 * 512 * 8          # algorithmic bytes of the noise kernel per frame (f32 in + f32 out)


def peaks():
    p = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d['hbm_gbs'], d['bf16_tflops'], d.get('bf16_tflops_sustained', d['bf16_tflops']), 'measured'
    return 6650.0, 1590.0, 1400.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU legs (the oracle port of the reference path; the only places bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------
def _cpu_noise_frames(args):
    model, nframes, seed = args
    import numpy as np
    from oracle import ref_numpy
    nm = ref_numpy.NoiseModelRef(model, include=4)
    np.random.seed(seed)
    y = np.random.rand(4, 512, 512).astype(np.float32)
    t0 = time.perf_counter()
    for _ in range(nframes):
        z = nm(y, params=SONY)
        z = np.maximum(np.minimum(z, 1.0), 0)
    return time.perf_counter() - t0


def cpu_noise_baseline(model, seconds=12.0):
    """Single-thread numpy port of noise.py:149-170 (+clip) on one core, bounded to ~`seconds`."""
    t1 = _cpu_noise_frames((model, 2, 2018))
    n = max(4, int(seconds / (t1 / 2)))
    t = _cpu_noise_frames((model, n, 2018))
    return {'value': n / t, 'unit': 'frames/s', 'cores': 1, 'kind': 'port',
            'sample': '%d frames of 4x512x512, numpy port of noise.py model %s, 1 thread' % (n, model)}


def reference_arm(a):
    """--impl reference: the reference's own CPU implementation of the path (the Python reference
    cannot be compiled or shipped; the numpy/torch oracle port restates it line by line), on all
    host cores: noise in a multiprocessing pool mirroring DataLoader(num_workers=nproc)
    (train_syn.py:78-80), U-Net fwd+L1+bwd+Adam with torch CPU on nproc threads."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    import numpy as np
    nproc = os.cpu_count() or 1
    workload = a.workload
    model = a.model
    per_step = max(1, nproc)              # one frame per worker per step
    t_noise = []
    with mp.Pool(nproc) as pool:
        for s in range(a.warmup + a.steps):
            t0 = time.perf_counter()
            pool.map(_cpu_noise_frames, [(model, 1, 1000 + s * nproc + i) for i in range(per_step)])
            dt = time.perf_counter() - t0
            if s >= a.warmup:
                t_noise.append(dt)
    noise_fps = per_step * len(t_noise) / sum(t_noise)
    out = {'impl': 'reference', 'metric': 'raw frames/sec (noise+U-Net)', 'unit': 'frames/s', 'n_gpus': a.gpus,
           'steps': a.steps, 'warmup': a.warmup, 'higher_is_better': True, 'scaling': 'weak',
           'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
           'config': config_dict(a, workload)}
    if workload == 'noise':
        fps = noise_fps
        sample = '%d frames/step of 4x512x512 in a %d-process pool, numpy port of noise.py model %s' % (per_step, nproc, model)
    else:
        import torch
        from oracle import unet_ref
        fps_unet, nthr = unet_ref.cpu_train_fps_best(steps=max(1, min(a.steps, 3)), warmup=1, batch=1)
        fps = 1.0 / (1.0 / noise_fps + 1.0 / fps_unet)
        sample = ('noise: %d frames/step in a %d-process pool; U-Net: torch CPU fp32 fwd+L1+bwd+Adam batch 1 x %d steps on %d threads '
                  '(best of min(nproc, 16/32/64); nproc = %d)'
                  % (per_step, nproc, max(1, min(a.steps, 3)), nthr, nproc))
    out.update({'value': fps, 'ms_per_step': 1000.0 / fps,
                'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': nproc, 'kind': 'port', 'sample': sample},
                'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                'gpu_launches': 0})
    print(json.dumps(out))


def config_dict(a, workload):
    if workload == 'noise':
        return {'workload': 'noise.py %s sampler, batch %d x 4x512x512 packed raw (SonyA7S2 params), f32 in/out' % (a.model, a.batch),
                'frames_per_step_per_gpu': a.batch, 'cache': 'inputs+outputs %d MiB per step > 126 MiB L2' % (a.batch * 8)}
    return {'workload': 'train_syn.py step: %s noise + U-Net fwd+L1+bwd+Adam, batch %d x 4x512x512 bf16, L1 loss (BASELINE configs[2])' % (a.model, a.batch),
            'global_batch': a.batch * a.gpus, 'frames_per_step_per_gpu': a.batch, 'parallelism': 'dp%d' % a.gpus,
            'cache': 'activations %s > 126 MiB L2' % 'of a step'}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--workload', default=None, choices=[None, 'noise', 'train'])
    ap.add_argument('--model', default='P+g', help="noise model string (reference semantics); G+P* = 'P+g'")
    ap.add_argument('--batch', type=int, default=None, help='frames per GPU per step')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    a = ap.parse_args()

    from eld_b200 import engine_available
    if a.workload is None:
        a.workload = 'train' if engine_available() else 'noise'
    if a.batch is None:
        a.batch = 8 if a.workload == 'train' else 32
    a.warmup = max(a.warmup, 3)

    if a.impl == 'reference':
        return reference_arm(a)

    import numpy as np
    import torch
    import torch.distributed as dist
    from eld_b200 import _lib
    from eld_b200.noise import NoiseModel

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    assert world == a.gpus, 'launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)' % (a.gpus, world)

    hbm_peak, tf_peak, tf_sus, peak_src = peaks()
    B = a.batch
    nm = NoiseModel(a.model, include=4, verbose=False, seed=2018)
    torch.manual_seed(2018 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None      # one nvidia-smi poller per job, not per rank
    extra = {}
    if a.workload == 'noise':
        # two alternating input/output sets so no step re-reads lines the previous one left in L2
        clean = [torch.rand(B, 4, 512, 512, device=dev) for _ in range(2)]
        noisy = [torch.empty_like(clean[0]) for _ in range(2)]
        host_in = torch.rand(B, 4, 512, 512).pin_memory()
        host_out = torch.empty(B, 4, 512, 512).pin_memory()
        dev_in = torch.empty(B, 4, 512, 512, device=dev)
        dev_out = torch.empty_like(dev_in)
        plist = [SONY] * B

        def step(i):
            nm.batch_gpu(clean[i & 1], params=plist, frame_id0=(i * world + rank) * B, out=noisy[i & 1])

        def step_e2e(i):
            dev_in.copy_(host_in, non_blocking=True)
            nm.batch_gpu(dev_in, params=plist, frame_id0=(i * world + rank) * B, out=dev_out)
            host_out.copy_(dev_out, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        h2d, d2h = host_in.numel() * 4, host_out.numel() * 4
        launches_per_step = 1
        kernel_bytes = B * FRAME_BYTES
        dtype = 'f32'
    else:
        from eld_b200.train_bench import make_train_steps
        step, step_e2e, h2d, d2h, launches_per_step, extra = make_train_steps(a, nm, dev, rank, world)
        kernel_bytes = None
        dtype = 'bf16'

    for i in range(a.warmup):
        step(i)
    barrier()
    l0 = _lib.launch_count(local)
    if sampler is not None:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for i in range(a.steps):
        step(a.warmup + i)
    ev[1].record()
    barrier()
    clocks = sampler.stop() if sampler is not None else None
    ms = ev[0].elapsed_time(ev[1])
    launches = _lib.launch_count(local) - l0
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    frames = B * world * a.steps
    value = frames / (ms * 1e-3)

    # end to end through the public API with HOST buffers (pinned), copies inside the timed region
    for i in range(2):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, a.steps // 2)
    for i in range(e2e_steps):
        step_e2e(i)
    barrier()
    te = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = B * world * e2e_steps / float(te.item())

    if a.workload == 'noise':
        ach = kernel_bytes * a.steps / (ms * 1e-3) / 1e9   # the step IS the kernel
        roof = {'bound': 'hbm', 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach / hbm_peak,
                'traffic': None, 'kernel': 'noise_packed_vec_kernel<%s>' % a.model, 'peak_source': peak_src,
                'algorithmic_bytes_per_launch': kernel_bytes}
    else:
        roof = extra.pop('roofline')
        roof['peak_source'] = peak_src

    out = {'metric': 'raw frames/sec (noise+U-Net)', 'value': value, 'unit': 'frames/s', 'n_gpus': world,
           'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': ms / a.steps, 'higher_is_better': True,
           'scaling': 'weak', 'vs_baseline': None, 'dtype': dtype, 'data': 'synthetic',
           'config': config_dict(a, a.workload), 'clocks': clocks,
           'e2e': {'value': e2e_value, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
           'gpu_launches': launches, 'roofline': roof}
    out.update(extra)
    if rank == 0:
        if not a.no_cpu_baseline and world == 1:      # the CPU baseline is timed at N = 1 only (the other ranks would idle)
            out['cpu_baseline'] = cpu_baseline(a)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(a):
    if a.workload == 'noise':
        return cpu_noise_baseline(a.model)
    import torch
    from oracle import unet_ref
    nb = cpu_noise_baseline(a.model, seconds=6.0)
    fps_unet, nthr = unet_ref.cpu_train_fps_best(steps=2, warmup=1, batch=1)
    fps = 1.0 / (1.0 / nb['value'] + 1.0 / fps_unet)
    return {'value': fps, 'unit': 'frames/s', 'cores': nthr, 'kind': 'port',
            'sample': nb['sample'] + '; U-Net torch CPU fp32 fwd+L1+bwd+Adam, batch 1 x 2 steps, %d threads - best of 16/32/64 (noise leg 1 thread)' % nthr}


if __name__ == '__main__':
    main()
