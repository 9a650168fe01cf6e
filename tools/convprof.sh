mkdir -p gpurun_out/prof
ELD_CONV_PROF=1 python eld_b200/smoke_unet.py > /dev/null 2>&1
ELD_CONV_PROF=1 python - > gpurun_out/prof/convprof.txt 2>&1 <<'PY'
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
ELD_CONV_PROF=1 ELD_CONV_DBG=25 python - > gpurun_out/prof/convprof_all3.txt 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, '.')
from eld_b200 import arch
torch.manual_seed(0)
net = arch.unet(4, 4).cuda()
x = torch.rand(8, 4, 512, 512, device='cuda'); t = torch.rand_like(x)
loss = torch.zeros((), device='cuda')
net.train_step(x, t, loss_out=loss)
torch.cuda.synchronize()
PY
tail -50 gpurun_out/prof/convprof.txt
