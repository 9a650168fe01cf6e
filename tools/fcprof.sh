#!/bin/bash
# in-kernel role / barrier-wait cycles of the two first-layer tiles (second step of a warm engine)
O=${1:-gpurun_out/prof}; mkdir -p $O
ELD_FC_PROF=1 python - > $O/fcprof.txt 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, '.')
from eld_b200 import arch
torch.manual_seed(0)
net = arch.unet(4, 4).cuda()
x = torch.rand(8, 4, 512, 512, device='cuda'); t = torch.rand_like(x)
loss = torch.zeros((), device='cuda')
for i in range(3):
    net.train_step(x, t, loss_out=loss)
    torch.cuda.synchronize()
PY
grep "first conv prof" $O/fcprof.txt | tail -4
