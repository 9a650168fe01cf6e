#!/usr/bin/env python
"""One training step (noise + U-Net fwd/bwd + Adam) inside a cudaProfilerStart/Stop window, for
    ncu --profile-from-start off ... python tools/ncu_step.py [batch]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eld_b200 import arch
from eld_b200.noise import NoiseModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
SONY = (2.2881136684755243, 6.4508722699636545, 15583, 208.9766365993794)
torch.manual_seed(2018)
net = arch.unet(4, 4).cuda()
opt = arch.FusedAdam(net, lr=1e-4, betas=(0.9, 0.999), weight_decay=0.0)
nm = NoiseModel('P+g', include=4, verbose=False, seed=2018)
clean = torch.rand(B, 4, 512, 512, device='cuda')
noisy = torch.empty_like(clean)
loss = torch.zeros((), device='cuda')


def step(i):
    nm.batch_gpu(clean, params=[SONY] * B, frame_id0=i * B, out=noisy)
    net.train_step(noisy, clean, loss_out=loss)
    opt.step()


for i in range(2):
    step(i)
torch.cuda.synchronize()
torch.cuda.profiler.start()
step(2)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('loss', float(loss))
