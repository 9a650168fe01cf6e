#!/usr/bin/env python
"""Summarise .ncu-rep files (ncu -i ... --page raw --csv) into the metrics the rooflines use, and
aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
    python tools/ncu_summary.py rep a.ncu-rep [b.ncu-rep ...]
    python tools/ncu_summary.py launches launches.csv
"""
import csv
import io
import re
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.sum', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size',
        'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active']


def rep(paths):
    for path in paths:
        out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(path, ': no data')
            continue
        head, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(head, r))
            print('== %s :: %s' % (path.split('/')[-1], d.get('Kernel Name', '?')[:100]))
            for k in KEYS:
                if k in d and d[k] not in ('', 'n/a'):
                    print('   %-78s %s %s' % (k, d[k], units[head.index(k)]))
            rd, wr, t = d.get('dram__bytes_read.sum'), d.get('dram__bytes_write.sum'), d.get('gpu__time_duration.sum')
            print()


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    head = rows[0]
    ik, iv, iu = head.index('Kernel Name'), head.index('Metric Value'), head.index('Metric Unit')
    agg, order, total = {}, [], 0.0
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(',', ''))
        except ValueError:
            continue
        scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(r[iu], 1e-3)
        name = re.sub(r'\(.*', '', r[ik])
        if name not in agg:
            agg[name] = [0, 0.0]
            order.append(name)
        agg[name][0] += 1
        agg[name][1] += v * scale
        total += v * scale
    print('launches: %d   total device time %.1f us (ncu: cold caches, serialised - compare SHARES, not absolutes)' % (sum(a[0] for a in agg.values()), total))
    for name in sorted(order, key=lambda n: -agg[n][1]):
        c, t = agg[name]
        print('%-64s n=%4d  %10.1f us  %5.1f%%' % (name[:64], c, t, 100 * t / total))


if __name__ == '__main__':
    (rep if sys.argv[1] == 'rep' else launches)(sys.argv[2:] if sys.argv[1] == 'rep' else sys.argv[2])
