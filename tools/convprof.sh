mkdir -p gpurun_out/prof
bash tools/quick.sh q3 > gpurun_out/prof/q3.out 2>&1
ELD_WGRAD_NOBIAS=1 python tools/profile_layers.py 8 > gpurun_out/prof/layers_nobias.txt 2>&1
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
cat gpurun_out/prof/q3.out; tail -9 gpurun_out/prof/layers_nobias.txt | head -3
grep -A200 "second step" gpurun_out/prof/convprof.txt | grep wgrad | cut -c1-400
