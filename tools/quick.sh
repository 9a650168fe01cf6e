# quick GPU check: parity tests + per-layer table.  usage: bash tools/quick.sh <tag>
O=gpurun_out/${1:-quick}; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > $O/pytest.txt
timeout 300 python tools/profile_layers.py 8 $O/layers.json > $O/layers.txt 2>&1
tail -4 $O/pytest.txt; head -1 $O/layers.txt; tail -9 $O/layers.txt
