"""Per-kernel time shares of ONE graph replay from an ncu launch list.

    ncu --metrics gpu__time_duration.sum[,sm__cycles_active.sum] --clock-control none --csv --log-file <csv> \
        python scripts/replay_loop.py 3 [pairs]
    python scripts/launch_shares.py <csv> [pairs]

With sm__cycles_active.sum in the list a second column reports the SM-time of a kernel family (sum over SMs of
their active cycles / 148 SMs / SM clock): what the family costs when OTHER forwards fill the idle SMs, i.e. in
the multi-stream regime the benchmark runs in.
"""
import csv
import re
import sys
from collections import OrderedDict, defaultdict

path = sys.argv[1]
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
launches = OrderedDict()                       # launch id -> dict(name, metrics)
for r in rows:
    d = launches.setdefault(int(r[0]), dict(name=r[4], m={}))
    d['m'][r[12]] = float(r[14].replace(',', ''))


def short(n):
    n = re.sub(r'\(.*', '', n.replace('void ', '').replace('<unnamed>::', '')).strip()
    n = re.sub(r'cub::DeviceRadixSortOnesweepKernel<.*', 'cub::DeviceRadixSortOnesweepKernel', n)
    n = re.sub(r'cub::DeviceRadixSortHistogramKernel<.*', 'cub::DeviceRadixSortHistogramKernel', n)
    n = re.sub(r'cub::DeviceRadixSortExclusiveSumKernel<.*', 'cub::DeviceRadixSortExclusiveSumKernel', n)
    n = re.sub(r'cub::DeviceScanKernel<.*', 'cub::DeviceScanKernel', n)
    n = re.sub(r'cub::DeviceScanInitKernel<.*', 'cub::DeviceScanInitKernel', n)
    n = re.sub(r'at::native::.*|at::.*', 'torch elementwise / cat', n)
    return n


names = [short(v['name']) for v in launches.values()]
vals = list(launches.values())
n = len(names)
period = next((p for p in range(50, n // 2 + 1) if names[n - p:] == names[n - 2 * p:n - p]), n)   # replays repeat exactly; a trimmed list holds one
agg = defaultdict(lambda: [0.0, 0, 0.0])
has_act = False
for nm, v in zip(names[n - period:], vals[n - period:]):
    a = agg[nm]
    a[0] += v['m'].get('gpu__time_duration.sum', 0.0) / 1e3          # ns -> us
    a[1] += 1
    if 'sm__cycles_active.sum' in v['m']:
        has_act = True
        a[2] += v['m']['sm__cycles_active.sum'] / 148.0 / 1965.0     # cycles -> us of a fully busy machine at 1965 MHz
tot = sum(v[0] for v in agg.values())
tot_act = sum(v[2] for v in agg.values())
out = sys.stdout
out.write(f'# one CUDA-graph replay of the forward ({pairs} pair(s) per forward), ncu --clock-control none\n')
out.write('# cold-cache serialised per-launch times: compare SHARES, not absolutes\n')
out.write(f'# launches {period}, sum {tot:.1f} us ({tot / pairs:.1f} us / pair)' +
          (f', SM-time {tot_act:.1f} us ({tot_act / pairs:.1f} us / pair)' if has_act else '') + '\n')
out.write('us,launches,share' + (',sm_time_us,sm_time_share' if has_act else '') + ',kernel\n')
for nm, (t, c, act) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    out.write(f'{t:.1f},{c},{t / tot:.3f}' + (f',{act:.1f},{act / max(tot_act, 1e-9):.3f}' if has_act else '') + f',"{nm}"\n')
