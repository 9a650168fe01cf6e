#!/bin/bash
# compute-sanitizer pass over one small training step + inference (the tiles rely on hand-rolled mbarrier
# protocols): usage: gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
O=gpurun_out/sanitize; mkdir -p $O
cat > /tmp/san_step.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from eld_b200 import arch
from eld_b200.noise import NoiseModel
torch.manual_seed(0)
net = arch.unet(4, 4).cuda()
opt = arch.FusedAdam(net, lr=1e-4)
nm = NoiseModel('P+g', include=4, verbose=False, seed=1)
clean = torch.rand(1, 4, 128, 256, device='cuda')
x = nm.batch_gpu(clean, params=(2.288, 6.45, 15583, 208.98), frame_id0=0)
out, loss = net.train_step(x, clean)
opt.step()
net.eval()
with torch.no_grad():
    y = net(torch.rand(1, 4, 48, 80, device='cuda'))
torch.cuda.synchronize()
print('step ok', float(loss), float(y.abs().mean()))
PY
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_step.py > $O/$tool.txt 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|step ok|RACECHECK SUMMARY|hazard" $O/$tool.txt | head -8
done
