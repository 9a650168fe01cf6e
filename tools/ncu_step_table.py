#!/usr/bin/env python
"""Per-launch table from `ncu --csv` (tools/ncu_round.sh step_sol.csv): duration, DRAM bytes, DRAM / L2 / tensor-pipe %,
joined with the engine's launch order (gpurun_out/.../layers.json) so every row carries its layer name.
    python tools/ncu_step_table.py step_sol.csv [layers.json] [traffic.json]
With a third argument the DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum per launch) is also written as
JSON - bench.py reads profiles/ncu_traffic.json for `roofline.traffic` (tensor tiles of a step) and
`roofline_noise.traffic`.
"""
import collections
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
head = rows[hi]
K, M, V = head.index('Kernel Name'), head.index('Metric Name'), head.index('Metric Value')
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= V:
        continue
    d = per.setdefault(int(r[0]), {'kernel': r[K]})
    d[r[M]] = r[V]
names = []
if len(sys.argv) > 2:
    names = [x['name'] for x in json.load(open(sys.argv[2]))['launches']]
ours = [d for d in per.values() if 'eld::' in d['kernel'] or d['kernel'].startswith('void noise') or 'kernel' in d['kernel']]


def f(d, k):
    try:
        return float(d.get(k, 'nan').replace(',', ''))
    except ValueError:
        return float('nan')


print('%-3s %-26s %-34s %9s %9s %9s %7s %7s %7s' % ('#', 'layer', 'kernel', 'us', 'rd MB', 'wr MB', 'dram%', 'L2%', 'tens%'))
tot = 0.0
i_name = 0
for i, d in enumerate(per.values()):
    k = d['kernel'].replace('void ', '').replace('eld::', '').split('(')[0][:34]
    lay = ''
    if names and not k.startswith('noise') and ('conv_umma' in k or 'wgrad' in k or 'maxpool' in k or 'head' in k or 'colsum' in k or 'pack_' in k or 'permute' in k or 'first_conv' in k):
        lay = names[i_name] if i_name < len(names) else ''
        i_name += 1
    us = f(d, 'gpu__time_duration.sum') / 1e3
    tot += us
    print('%-3d %-26s %-34s %9.1f %9.1f %9.1f %7.1f %7.1f %7.1f' % (
        i, lay, k, us, f(d, 'dram__bytes_read.sum') / 1e6, f(d, 'dram__bytes_write.sum') / 1e6,
        f(d, 'DRAM Throughput'), f(d, 'L2 Cache Throughput'), f(d, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed')))
print('total %.1f us over %d launches (ncu: serialised, cold caches, --clock-control none)' % (tot, len(per)))
if len(sys.argv) > 3:
    import os
    tiles = noise = 0.0
    rows_out = []
    for d in per.values():
        k = d['kernel']
        by = f(d, 'dram__bytes_read.sum') + f(d, 'dram__bytes_write.sum')
        rows_out.append({'kernel': k.replace('void ', '').split('(')[0][:60], 'dram_bytes': by, 'us': f(d, 'gpu__time_duration.sum') / 1e3})
        if 'conv_umma_kernel' in k or 'wgrad_conv_kernel' in k or 'wgrad_umma_kernel' in k or 'first_conv' in k:
            tiles += by
        if 'noise_packed' in k:
            noise = by
    out = {}
    if os.path.exists(sys.argv[3]):
        out = json.load(open(sys.argv[3]))
    out['train'] = {'tensor_tile_bytes_per_step': tiles, 'noise_bytes_per_launch': noise, 'launches': rows_out,
                    'source': 'ncu --clock-control none, SpeedOfLight + MemoryWorkloadAnalysis sections, one training step (tools/ncu_step.py)'}
    json.dump(out, open(sys.argv[3], 'w'), indent=0)
