#!/usr/bin/env python
"""Per-launch table of one training step (CUDA events recorded by the engine, eld_unet_profile).
    python tools/profile_layers.py [batch] [out.json]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eld_b200 import arch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(2018)
net = arch.unet(4, 4).cuda()
x = torch.rand(B, 4, 512, 512, device='cuda')
t = torch.rand(B, 4, 512, 512, device='cuda')
recs = net.profile(x, t, steps=5)
tot = sum(r['ms'] for r in recs)
print('batch %d: sum of kernel times %.3f ms  -> %.1f frames/s (U-Net fwd+bwd only)' % (B, tot, B / tot * 1e3))
agg = {}
for r in recs:
    tf = r['flops'] / r['ms'] / 1e9 if r['ms'] > 0 else 0
    gb = r['bytes'] / r['ms'] / 1e6 if r['ms'] > 0 else 0
    print('%-24s %8.3f ms %5.1f%% %8.1f TF/s %8.1f GB/s' % (r['name'], r['ms'], 100 * r['ms'] / tot, tf, gb))
    k = r['name'].split('.')[1]
    a = agg.setdefault(k, [0.0, 0.0])
    a[0] += r['ms']; a[1] += r['flops']
for k, (ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print('  %-14s %8.3f ms %5.1f%%  %8.1f TF/s' % (k, ms, 100 * ms / tot, fl / ms / 1e9 if ms else 0))
if len(sys.argv) > 2:
    json.dump({'batch': B, 'total_ms': tot, 'launches': recs}, open(sys.argv[2], 'w'), indent=0)
