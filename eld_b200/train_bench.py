"""bench.py's training workload (BASELINE configs[2]): noise kernel -> U-Net fwd + L1 + bwd ->
(NCCL all-reduce of the flat gradient) -> fused Adam, one process per GPU."""
import json
import os

import torch
import torch.distributed as dist

from . import arch

SONY = (2.2881136684755243, 6.4508722699636545, 15583, 208.9766365993794)


def make_train_steps(a, nm, dev, rank, world):
    B = a.batch
    torch.manual_seed(2018)                       # same init on every rank (reference default --seed 2018)
    net = arch.unet(4, 4).to(dev)
    opt = arch.FusedAdam(net, lr=1e-4, betas=(0.9, 0.999), weight_decay=0.0)
    torch.manual_seed(2018 + rank)
    # two alternating clean batches; a step touches ~2.7 GB of activations so nothing survives in L2
    clean = [torch.rand(B, 4, 512, 512, device=dev) for _ in range(2)]
    noisy = torch.empty_like(clean[0])
    loss = torch.zeros((), device=dev)
    plist = [SONY] * B
    host_clean = torch.rand(B, 4, 512, 512).pin_memory()
    dev_clean = torch.empty(B, 4, 512, 512, device=dev)
    host_loss = torch.zeros(1).pin_memory()

    def body(target, i):
        nm.batch_gpu(target, params=plist, frame_id0=(i * world + rank) * B, out=noisy)
        net.train_step(noisy, target, loss_out=loss)
        if world > 1:
            dist.all_reduce(net.flat_grads)
        opt.step(grad_scale=1.0 / world)

    def step(i):
        body(clean[i & 1], i)

    # end to end: every step's clean batch comes from pinned host memory.  Like a DataLoader with
    # pin_memory + prefetch (train_syn.py:78-80) the copy of batch i+1 runs on a side stream while step i
    # computes; every copy and the D2H of the loss are inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    dev_bufs = [torch.empty(B, 4, 512, 512, device=dev) for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    state = {'next': None}

    def issue_copy(i):
        k = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[k])            # the step that last read this buffer is done
            dev_bufs[k].copy_(host_clean, non_blocking=True)
            copied[k].record(copy_stream)

    def step_e2e(i):
        k = i & 1
        if state['next'] != i:
            issue_copy(i)
        cur = torch.cuda.current_stream()
        cur.wait_event(copied[k])
        body(dev_bufs[k], i)
        consumed[k].record(cur)
        issue_copy(i + 1)
        state['next'] = i + 1
        host_loss.copy_(loss.reshape(1), non_blocking=True)
        cur.synchronize()

    # ---- live per-launch profile for the roofline entry (a separate, untimed pass) ---------------
    extra = {}
    nm.batch_gpu(clean[0], params=plist, frame_id0=0, out=noisy)
    recs = net.profile(noisy, clean[0], steps=3)
    tens = [r for r in recs if r['name'].split('.')[1] in ('fprop', 'dgrad', 'wgrad') and not r['name'].startswith(('conv1_1', 'conv10'))]
    t_ms = sum(r['ms'] for r in tens)
    fl = sum(r['flops'] for r in tens)
    total_ms = sum(r['ms'] for r in recs)
    import bench as _b
    hbm_peak, tf_peak, tf_sus, _src = _b.peaks()
    ach = fl / (t_ms * 1e-3) / 1e12
    extra['roofline'] = {'bound': 'tensor', 'achieved': ach, 'peak': tf_sus, 'unit': 'TFLOP/s', 'frac': ach / tf_sus,
                         'traffic': None, 'kernel': 'conv_umma_kernel + wgrad_umma_kernel (all %d tcgen05 launches of a step)' % len(tens),
                         'algorithmic_flops_per_step': fl, 'tensor_ms_per_step': t_ms, 'all_kernels_ms_per_step': total_ms,
                         'share_of_step': t_ms / total_ms, 'peak_kind': 'bf16_tflops_sustained (kernels timed inside a long step)'}
    if rank == 0:
        os.makedirs(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out'), exist_ok=True)
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', 'layer_profile.json')
        with open(path, 'w') as f:
            json.dump(recs, f, indent=0)
    launches = len(recs) + 1 + 1 + 2      # + noise + adam + 2 memsets are not kernels of ours; counted by the ctx anyway
    return step, step_e2e, host_clean.numel() * 4, 4, launches, extra
