#!/bin/bash
# in-kernel role / barrier-wait cycles of every tcgen05 launch of one training step (second step of a warm engine)
O=${1:-gpurun_out/prof}; mkdir -p $O
ELD_CONV_PROF=1 python - > $O/convprof.txt 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, '.')
from eld_b200 import arch
torch.manual_seed(0)
net = arch.unet(4, 4).cuda()
x = torch.rand(8, 4, 512, 512, device='cuda'); t = torch.rand_like(x)
loss = torch.zeros((), device='cuda')
net.train_step(x, t, loss_out=loss)
torch.cuda.synchronize()
print('----- second step', file=sys.stderr, flush=True)
net.train_step(x, t, loss_out=loss)
torch.cuda.synchronize()
PY
grep -c prof $O/convprof.txt
