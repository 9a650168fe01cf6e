#!/bin/bash
# One GPU-box visit: parity tests, bench lines for every BASELINE config, per-launch table, ncu launch list.
# usage: gpurun --timeout 2400 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-run}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_parity_fullsize_gpu.py 2>&1 | tail -25 ) > $O/pytest.txt 2>&1
( time timeout 900 python -m pytest tests/test_parity_fullsize_gpu.py -m gpu -q -s 2>&1 | tail -120 ) > $O/pytest_fullsize.txt 2>&1
timeout 600 python bench.py > $O/bench_train.json 2> $O/bench_train.err
timeout 300 python bench.py --workload infer > $O/bench_infer.json 2> $O/bench_infer.err
timeout 300 python bench.py --workload noise --model P+g > $O/bench_noise_P+g.json 2> $O/bench_noise.err
timeout 300 python bench.py --workload noise --model p+g --no-cpu-baseline > $O/bench_noise_p+g.json 2>> $O/bench_noise.err
timeout 300 python bench.py --workload noise --model ELD:P+G+B+R+U --batch 4 --no-cpu-baseline > $O/bench_noise_full.json 2>> $O/bench_noise.err
timeout 300 python bench.py --workload fullframe > $O/bench_fullframe.json 2> $O/bench_fullframe.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
timeout 300 python tools/profile_layers.py 8 $O/layers.json > $O/layers.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-onbox > $O/bench_under_ncu.log 2>&1
tail -4 $O/pytest.txt; tail -8 $O/pytest_fullsize.txt; for f in $O/bench_*.json; do echo "== $f"; cut -c1-1500 $f; echo; done; tail -n 5 $O/bench_*.err; tail -12 $O/layers.txt; tail -3 $O/smoke.txt
