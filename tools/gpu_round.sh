#!/bin/bash
# One GPU-box visit: parity tests, bench lines, per-launch table, ncu launch list.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-run}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) > $O/pytest.txt 2>&1
timeout 600 python bench.py > $O/bench_train.json 2> $O/bench_train.err
timeout 300 python bench.py --workload noise --model P+g > $O/bench_noise_P+g.json 2> $O/bench_noise.err
timeout 300 python bench.py --workload noise --model p+g --no-cpu-baseline > $O/bench_noise_p+g.json 2>> $O/bench_noise.err
timeout 300 python tools/profile_layers.py 8 $O/layers.json > $O/layers.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
tail -3 $O/pytest.txt; cat $O/bench_train.json; cat $O/bench_noise_P+g.json; cat $O/bench_noise_p+g.json; tail -12 $O/layers.txt
