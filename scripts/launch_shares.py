"""Per-kernel time shares of ONE graph replay from an ncu launch list.

    ncu --metrics gpu__time_duration.sum[,sm__cycles_active.sum] --clock-control none --csv --log-file <csv> \
        python scripts/replay_loop.py 3 [pairs]
    python scripts/launch_shares.py <csv> [pairs] [traffic.json]

With dram__bytes_read.sum / dram__bytes_write.sum in the list and a third argument, the DRAM traffic of the kernel
families (per launch and per replay) is merged into traffic.json under str(pairs): bench.py's `roofline.traffic`.

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
SCALE = {'byte': 1.0, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'nsecond': 1.0, 'ns': 1.0, 'usecond': 1e3, 'us': 1e3,
         'msecond': 1e6, 'ms': 1e6, 'second': 1e9}
for r in rows:
    d = launches.setdefault(int(r[0]), dict(name=r[4], m={}))
    d['m'][r[12]] = float(r[14].replace(',', '')) * SCALE.get(r[13].strip().lower(), 1.0)      # bytes, ns, cycles


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

if len(sys.argv) > 3:
    import json
    FAMILIES = [('kpconv_gather', r'k_kpconv_agg|k_kpconv_c1'), ('gemm_tf32x3', r'k_gemm_tf32x3|k_splitk_reduce'),
                ('attention_core', r'k_mha_')]
    fam = {}
    for nm, v in zip(names[n - period:], vals[n - period:]):
        key = next((f for f, pat in FAMILIES if re.search(pat, nm)), 'other')
        d = fam.setdefault(key, dict(launches=0, dram_bytes_per_replay=0.0, us_per_replay_under_ncu=0.0))
        d['launches'] += 1
        d['dram_bytes_per_replay'] += v['m'].get('dram__bytes_read.sum', 0.0) + v['m'].get('dram__bytes_write.sum', 0.0)
        d['us_per_replay_under_ncu'] += v['m'].get('gpu__time_duration.sum', 0.0) / 1e3
    for d in fam.values():
        d['dram_bytes_per_launch'] = d['dram_bytes_per_replay'] / max(d['launches'], 1)
    tb = sum(d['dram_bytes_per_replay'] for d in fam.values())
    fam['whole_forward'] = dict(launches=period, dram_bytes_per_replay=tb, dram_bytes_per_launch=tb,
                                us_per_replay_under_ncu=tot)
    try:
        allt = json.load(open(sys.argv[3]))
    except Exception:
        allt = {}
    allt[str(pairs)] = fam
    allt['_source'] = ('scripts/launch_shares.py over the ncu launch lists profiles/r02_launches_b*.csv (dram__bytes_read.sum + '
                       'dram__bytes_write.sum of every kernel of one graph replay)')
    json.dump(allt, open(sys.argv[3], 'w'), indent=1, sort_keys=True)
