#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel of libeld_b200.so the counts of the Blackwell-specific instructions
(tcgen05 = UTCHMMA / UTCBAR / LDTM / UTCATOMSWS..., TMA = UTMALDG / UBLKCP / UTMAPF, mbarrier = SYNCS) and the totals.
    python tools/sass_summary.py > profiles/r02_sass_summary.txt        (runs here: cuobjdump needs no GPU)"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'eld_b200', 'libeld_b200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
KEYS = ['UTCHMMA', 'UTCBAR', 'LDTM', 'UTCATOMSWS', 'UTMALDG', 'UBLKCP', 'UTMACCTL', 'SYNCS', 'ELECT', 'REDG', 'ATOMG', 'LDGSTS', 'MUFU']
cur, funcs = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        funcs[cur] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,5}\*/\s+(.*?);', line)
    if m and cur:
        t = m.group(1).split()
        op = t[1] if t[0].startswith('@') and len(t) > 1 else t[0]
        funcs[cur][op.split('.')[0]] += 1
        funcs[cur]['_total'] += 1
tot = collections.Counter()
dem = lambda n: subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip().split('(')[0][:70]
print('%-72s %7s  %s' % ('kernel', 'instrs', '  '.join('%s' % k for k in KEYS)))
for f, c in funcs.items():
    for k in KEYS:
        tot[k] += c[k]
    print('%-72s %7d  %s' % (dem(f), c['_total'], '  '.join('%*d' % (len(k), c[k]) for k in KEYS)))
print('%-72s %7s  %s' % ('TOTAL', '', '  '.join('%*d' % (len(k), tot[k]) for k in KEYS)))
