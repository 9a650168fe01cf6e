#!/bin/bash
# A/B of the gradient communicator's CTA limit: usage: gpurun --gpus N --timeout 900 -- 'bash tools/ddp_ctas.sh N'
N=${1:-2}
O=gpurun_out/ddpctas$N; mkdir -p $O
export NCCL_DEBUG=WARN
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-onbox > $O/bench_1gpu.json 2> $O/err.txt
port=29700
for c in 0 2 4 8 16; do
  port=$((port+1))
  ELD_NCCL_MAX_CTAS=$c timeout 400 $TR --master-port $port bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_ctas$c.json 2>> $O/err.txt
done
python - <<PY
import json
one = json.loads(open('$O/bench_1gpu.json').read().strip().splitlines()[-1])
print('1 GPU: %.3f ms/step  %.1f frames/s' % (one['ms_per_step'], one['value']))
for c in (0, 2, 4, 8, 16):
    try:
        d = json.loads(open('$O/bench_ctas%d.json' % c).read().strip().splitlines()[-1])
        print('max_ctas %2s: %.3f ms/step  %.1f frames/s  efficiency %.3f' % (c or 'default', d['ms_per_step'], d['value'], d['value'] / ($N * one['value'])))
    except Exception as e:
        print('max_ctas', c, 'failed', e)
PY
grep -v "^\*\|OMP_NUM\|^$" $O/err.txt | tail -4
