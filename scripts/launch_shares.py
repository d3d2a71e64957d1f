"""Per-kernel time shares of ONE graph replay from an ncu launch list
(ncu --metrics gpu__time_duration.sum --csv --log-file <csv> python scripts/replay_loop.py 3)."""
import csv, re, sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
names = [re.sub(r'\(.*', '', r[4].replace('void ', '').replace('<unnamed>::', '')).strip() for r in rows]
full = [re.sub(r'^void ', '', r[4].replace('<unnamed>::', '')) for r in rows]
ns = [float(r[14]) for r in rows]
n = len(names)
period = next(p for p in range(50, n // 2) if names[n - p:] == names[n - 2 * p:n - p])   # replays repeat exactly
tail_names = [re.sub(r'\(.*', '', f) for f in full[n - period:]]
agg = defaultdict(lambda: [0.0, 0])
for nm, t in zip(tail_names, ns[n - period:]):
    agg[nm][0] += t / 1e3
    agg[nm][1] += 1
tot = sum(v[0] for v in agg.values())
out = sys.stdout
out.write('# one CUDA-graph replay of the forward (BASELINE config 2, 1 pair), ncu --metrics gpu__time_duration.sum --clock-control none\n')
out.write('# cold-cache serialised per-launch times: compare SHARES, not absolutes\n')
out.write(f'# launches {period}, sum {tot:.1f} us\n')
out.write('us,launches,share,kernel\n')
for nm, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    out.write(f'{t:.1f},{c},{t / tot:.3f},"{nm}"\n')
