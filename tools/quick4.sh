#!/bin/bash
TAG=${1:-q}
O=gpurun_out/$TAG; mkdir -p $O
( timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_model_gpu.py tests/test_isp_gpu.py -m gpu -q 2>&1 | tail -15 ) > $O/pytest.txt 2>&1
timeout 400 python bench.py --no-cpu-baseline --no-onbox > $O/bench_train.json 2> $O/bench_train.err
tail -6 $O/pytest.txt; cut -c1-400 $O/bench_train.json
