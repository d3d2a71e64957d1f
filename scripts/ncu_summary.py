"""Summaries of one `ncu --set full` capture of a graph replay (scripts/make_profiles.sh) for profiles/.

    ncu -i <rep> --page raw --csv > raw.csv
    python scripts/ncu_summary.py raw.csv <pairs> <out prefix> [traffic.json]

Writes <out prefix>_<family>_summary.csv (one row per launch, the columns a reader needs to judge the kernel:
duration, DRAM bytes, L2/L1 hit rates, issue / warp / tensor-pipe activity, stall ratios, occupancy limits) for the
families kpconv_gather, gemm_tf32x3, attention_core and other, and merges the per-launch DRAM traffic of each family
(and of the whole replay) into traffic.json under the key str(pairs): bench.py's `roofline.traffic`.
"""
import csv
import json
import re
import sys

raw, pairs, prefix = sys.argv[1], sys.argv[2], sys.argv[3]
traffic_path = sys.argv[4] if len(sys.argv) > 4 else None
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
KEEP = ['ID', 'Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']
keep = [k for k in KEEP if k in col]
FAMILIES = [('kpconv_gather', r'k_kpconv_agg|k_kpconv_c1'), ('gemm_tf32x3', r'k_gemm_tf32x3|k_splitk_reduce'),
            ('attention_core', r'k_mha_')]


def family(name):
    for fam, pat in FAMILIES:
        if re.search(pat, name):
            return fam
    return 'other'


def fnum(r, key):
    try:
        return float(r[col[key]].replace(',', ''))
    except Exception:
        return 0.0


def to_bytes(r, key):
    u = units[col[key]].lower()
    scale = {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)
    return fnum(r, key) * scale


by = {}
for r in data:
    by.setdefault(family(r[col['Kernel Name']]), []).append(r)
note = (f'# ncu --set full --clock-control none --profile-from-start off, ONE CUDA-graph replay of BASELINE config 2 with '
        f'{pairs} pair(s) per forward (scripts/replay_loop.py 3 {pairs} 2 - profile); capacity-shaped launches, cold '
        f'caches, serialised kernels')
traffic = {}
tot_t = tot_b = 0.0
for fam, rs in by.items():
    with open(f'{prefix}_{fam}_summary.csv', 'w', newline='') as f:
        f.write(f'"{note}"\n')
        w = csv.writer(f)
        w.writerow(keep)
        w.writerow([units[col[k]] for k in keep])
        for r in rs:
            w.writerow([re.sub(r'\(.*', '', r[col[k]]).replace('void ', '').replace('<unnamed>::', '')
                        if k == 'Kernel Name' else r[col[k]] for k in keep])
    b = sum(to_bytes(r, 'dram__bytes_read.sum') + to_bytes(r, 'dram__bytes_write.sum') for r in rs)
    t = sum(fnum(r, 'gpu__time_duration.sum') for r in rs)
    tot_b += b; tot_t += t
    traffic[fam] = dict(launches=len(rs), dram_bytes_per_launch=b / len(rs), dram_bytes_per_replay=b,
                        us_per_replay_under_ncu=t)
traffic['whole_forward'] = dict(launches=len(data), dram_bytes_per_launch=tot_b, dram_bytes_per_replay=tot_b,
                                us_per_replay_under_ncu=tot_t)
for fam, d in traffic.items():
    print(f'{fam:16s} launches {d["launches"]:4d}  DRAM {d["dram_bytes_per_replay"] / 1e6:9.1f} MB / replay  '
          f'{d["us_per_replay_under_ncu"]:8.1f} us under ncu')
if traffic_path:
    try:
        allt = json.load(open(traffic_path))
    except Exception:
        allt = {}
    allt[str(pairs)] = traffic
    allt['_source'] = 'scripts/ncu_summary.py over ncu --set full captures of one graph replay (see profiles/README.md)'
    json.dump(allt, open(traffic_path, 'w'), indent=1, sort_keys=True)
