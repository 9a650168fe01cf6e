#!/bin/bash
# A/B visit: usage: gpurun --timeout 900 -- 'bash tools/quick5.sh tag "VAR1=1" "VAR2=32" ...'   (each extra arg = one B arm)
TAG=${1:-q}; shift
O=gpurun_out/$TAG; mkdir -p $O
( timeout 600 python -m pytest tests/test_parity_fullsize_gpu.py tests/test_unet_gpu.py tests/test_model_gpu.py tests/test_conv_gpu.py -m gpu -q 2>&1 | tail -15 ) > $O/pytest.txt 2>&1
timeout 300 python tools/profile_layers.py 8 $O/layers.json > $O/layers.txt 2>&1
timeout 400 python bench.py --no-cpu-baseline --no-onbox > $O/bench_train.json 2> $O/bench_train.err
tail -6 $O/pytest.txt; echo "== A"; grep -E "sum of" $O/layers.txt; tail -9 $O/layers.txt; cut -c1-330 $O/bench_train.json; echo
i=0
for arm in "$@"; do
  i=$((i+1))
  env $arm timeout 300 python tools/profile_layers.py 8 > $O/layers_B$i.txt 2>&1
  env $arm timeout 400 python bench.py --no-cpu-baseline --no-onbox > $O/bench_train_B$i.json 2>> $O/bench_train.err
  echo "== B$i ($arm)"; grep -E "sum of" $O/layers_B$i.txt; tail -9 $O/layers_B$i.txt; cut -c1-330 $O/bench_train_B$i.json; echo
done
