"""Per-kernel counts of the SASS mnemonics that prove the Blackwell path (B200_PROFILING.md):
UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTMALDG (TMA), HMMA (legacy mma.sync), LDGSTS (cp.async).

    python scripts/sass_summary.py [regtr_b200/libregtr_b200.so] > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          'regtr_b200', 'libregtr_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
pat = re.compile(r'\b(UTC[A-Z]*MMA|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|HMMA|LDGSTS|UTCBAR|SYNCS|RED|ATOM)\b')
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        k = re.search(r'(k_\w+(?:<[^>(]*>)?)', name)
        name = k.group(1) if k else re.sub(r'\(.*', '', name)[:70]
        cur = per.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    m = pat.search(line)
    if m and '/*' in line:
        cur[m.group(1)] += 1
cols = ['UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTCBAR', 'HMMA', 'LDGSTS', 'RED', 'ATOM']
print('# cuobjdump -sass of', os.path.basename(lib), '(sm_100a): instruction counts per kernel')
print('# UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA load, UTCBAR = tcgen05.commit,')
print('# HMMA = legacy mma.sync, LDGSTS = cp.async, RED / ATOM = global atomics')
print(f'{"kernel":70s} ' + ' '.join(f'{c:>8s}' for c in cols))
tot = collections.Counter()
for name, cnt in per.items():
    if not any(cnt.values()):
        continue
    print(f'{name[:70]:70s} ' + ' '.join(f'{cnt.get(c, 0):8d}' for c in cols))
    tot.update(cnt)
print(f'{"TOTAL":70s} ' + ' '.join(f'{tot.get(c, 0):8d}' for c in cols))
