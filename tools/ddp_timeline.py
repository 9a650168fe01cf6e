#!/usr/bin/env python
"""Where the gradient all-reduce sits inside a data-parallel training step (CUDA events on the device, rank 0):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29555 tools/ddp_timeline.py
Prints, per bucket, the interval [ready, reduced] relative to the step start next to the end of backward - a bucket whose
`reduced` time precedes `backward_end` cost nothing; only the tail of the last (encoder) bucket is exposed."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from eld_b200 import arch

local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
world, rank = dist.get_world_size(), dist.get_rank()
torch.manual_seed(2018)
net = arch.unet(4, 4).cuda()
opt = arch.FusedAdam(net, lr=1e-4)
x = torch.rand(8, 4, 512, 512, device='cuda')
t = torch.rand(8, 4, 512, 512, device='cuda')
# correctness first: the bucketed, overlapped exchange must equal one blocking all-reduce of the locally computed gradient
torch.manual_seed(100 + rank)
xr = torch.rand(8, 4, 512, 512, device='cuda')
net.train_step(xr, t)
want = net.flat_grads.clone()
dist.all_reduce(want)
net.train_step_ddp(xr, t)
net.join_allreduce()
torch.cuda.synchronize()
got = net.flat_grads.clone()
rel = ((got - want).norm() / want.norm()).item()
peers = [torch.empty_like(got) for _ in range(world)]
dist.all_gather(peers, got)
same = all(torch.equal(peers[0], q) for q in peers)
if rank == 0:
    print('bucketed all-reduce vs blocking all-reduce of the same local gradients: rel-L2 %.2e (fp32 atomics reorder the local sums); '
          'identical on all %d ranks: %s' % (rel, world, same))
assert rel < 1e-3 and same
for _ in range(5):
    net.train_step_ddp(x, t)
    opt.step(grad_scale=1.0 / world)
rows = []
for _ in range(5):
    tl = {}
    net.train_step_ddp(x, t, timeline=tl)
    opt.step(grad_scale=1.0 / world)
    end = torch.cuda.Event(enable_timing=True); end.record()
    torch.cuda.synchronize()
    s0 = tl['step_start']
    rows.append({'backward_end': s0.elapsed_time(tl['backward_end']), 'joined': s0.elapsed_time(tl['allreduce_joined']),
                 'adam_end': s0.elapsed_time(end),
                 'buckets': [(s0.elapsed_time(a), s0.elapsed_time(b), nbytes) for a, b, nbytes in tl['buckets']]})
if rank == 0:
    names = ['decoder upv6..conv10_1', 'bottleneck conv5_*', 'encoder conv2_1..conv4_2', 'first layer conv1_1..conv1_2']
    print('world %d, batch 8 x 4x512x512 per GPU; times in ms from the step start (median of 5 steps)' % world)
    med = lambda v: sorted(v)[len(v) // 2]
    print('backward_end %.3f   all-reduce joined %.3f   adam_end %.3f' % (med([r['backward_end'] for r in rows]),
          med([r['joined'] for r in rows]), med([r['adam_end'] for r in rows])))
    for k in range(len(rows[0]['buckets'])):
        a = med([r['buckets'][k][0] for r in rows]); b = med([r['buckets'][k][1] for r in rows])
        print('bucket %d %-26s %5.1f MB  ready %.3f  reduced %.3f  (%.0f us)' % (k, names[k], rows[0]['buckets'][k][2] / 1e6, a, b, (b - a) * 1e3))
dist.destroy_process_group()
