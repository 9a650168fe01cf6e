#!/bin/bash
# usage: gpurun --timeout 900 -- 'bash tools/quick3.sh tag'
TAG=${1:-q}
O=gpurun_out/$TAG; mkdir -p $O
( timeout 600 python -m pytest tests/test_parity_fullsize_gpu.py tests/test_unet_gpu.py tests/test_model_gpu.py tests/test_conv_gpu.py -m gpu -q 2>&1 | tail -25 ) > $O/pytest.txt 2>&1
timeout 300 python tools/profile_layers.py 8 $O/layers.json > $O/layers.txt 2>&1
timeout 400 python bench.py --no-cpu-baseline --no-onbox > $O/bench_train.json 2> $O/bench_train.err
ELD_OVERLAP=1 timeout 400 python bench.py --no-cpu-baseline --no-onbox > $O/bench_train_overlap.json 2>> $O/bench_train.err
ELD_SKIP=1 timeout 300 python - > $O/convprof.txt 2>&1 <<'PY'
import torch, sys, os
if os.environ.get("ELD_SKIP"): sys.exit(0)
sys.path.insert(0, '.')
from eld_b200 import arch
torch.manual_seed(0)
net = arch.unet(4, 4).cuda()
x = torch.rand(8, 4, 512, 512, device='cuda'); t = torch.rand_like(x)
loss = torch.zeros((), device='cuda')
for i in range(3):
    print('----- step', i, file=sys.stderr, flush=True)
    net.train_step(x, t, loss_out=loss)
    torch.cuda.synchronize()
PY
tail -25 $O/pytest.txt; cat $O/bench_train.json; grep -E "sum of|conv1_1|conv1_2|input" $O/layers.txt; cat $O/bench_train_overlap.json
