mkdir -p gpurun_out/prof
for d in 0 8; do
ELD_CONV_DBG=$d ELD_WGRAD_NOBIAS=1 ELD_CONV_PROF=1 python - > gpurun_out/prof/convprof_dbg$d.txt 2>&1 <<'PY'
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
echo "== dbg $d"; grep -A200 "second step" gpurun_out/prof/convprof_dbg$d.txt | grep wgrad | cut -c40-130,215-400 | head -9
done
