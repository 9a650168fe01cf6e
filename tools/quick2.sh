#!/bin/bash
# quick validation + A/B visit: usage: gpurun --timeout 900 -- 'bash tools/quick2.sh tag'
TAG=${1:-q}
O=gpurun_out/$TAG; mkdir -p $O
( timeout 600 python -m pytest tests/test_parity_fullsize_gpu.py tests/test_unet_gpu.py tests/test_model_gpu.py tests/test_conv_gpu.py -m gpu -q 2>&1 | tail -15 ) > $O/pytest.txt 2>&1
timeout 300 python tools/profile_layers.py 8 $O/layers.json > $O/layers.txt 2>&1
timeout 400 python bench.py --no-cpu-baseline --no-onbox > $O/bench_train.json 2> $O/bench_train.err
for g in 32 128; do ELD_L2_FETCH=$g timeout 300 python tools/profile_layers.py 8 > $O/layers_l2fetch$g.txt 2>&1; done
tail -5 $O/pytest.txt; cat $O/bench_train.json; grep -E "sum of|conv10|pool.bwd|upv9|upv8" $O/layers.txt; for g in 32 128; do echo "== L2 fetch $g"; grep -E "sum of|pool.bwd|upv9" $O/layers_l2fetch$g.txt; done
