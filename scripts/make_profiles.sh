#!/bin/bash
# GPU side of the profiles/ refresh (run through gpurun from the repo root):
#   scripts/gpu.sh 2400 gpurun_out/prof.log 'bash scripts/make_profiles.sh'
# bench lines (never under ncu), launch lists and one full ncu capture of a graph replay per batch size
# (`lite`: bench lines and launch lists only; `families`: + full captures of the gather / GEMM / attention kernels;
# `full`: every kernel -- ~20 GPU-minutes per batch size).
mode=${1:-full}
out=gpurun_out/final; mkdir -p $out
python bench.py > $out/bench_n1.json 2> $out/bench_n1.err
python bench.py --pairs 8 --cpu-baseline 0 > $out/bench_n1_pairs8.json 2> $out/bench_n1_pairs8.err
python bench.py --config 3 --attention bf16_tc --cpu-baseline 0 > $out/bench_n1_config3_bf16tc.json 2> $out/bench_n1_config3_bf16tc.err
python bench.py --attention tf32_tc --cpu-baseline 0 > $out/bench_n1_tf32tc.json 2> $out/bench_n1_tf32tc.err
python bench.py --config 4 --cpu-baseline 0 > $out/bench_n1_config4_share.json 2> $out/bench_n1_config4_share.err
python bench.py --config 5 --cpu-baseline 0 > $out/bench_n1_config5_share.json 2> $out/bench_n1_config5_share.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
for b in 1 8; do
  ncu --metrics gpu__time_duration.sum,sm__cycles_active.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      --csv --log-file $out/launches_b$b.csv python scripts/replay_loop.py 3 $b > $out/rl_b$b.log 2>&1
  [ "$mode" = lite ] && continue
  fam=""; [ "$mode" = families ] && fam="-k regex:k_kpconv_agg|k_kpconv_c1|k_gemm_tf32x3|k_mha_"   # the three roofline families only
  ncu --set full --clock-control none --profile-from-start off $fam -o $out/prof_b$b -f \
      python scripts/replay_loop.py 3 $b 2 - profile > $out/ncu_b$b.log 2>&1
  ncu -i $out/prof_b$b.ncu-rep --page raw --csv > $out/raw_b$b.csv 2>/dev/null
done
[ "$mode" != full ] && { rm -f $out/prof_b1.ncu-rep $out/prof_b8.ncu-rep; ls -la $out; exit 0; }
for a in tf32_tc bf16_tc; do
  ncu --set full --clock-control none --profile-from-start off -k regex:k_mha -o $out/prof_b8_$a -f \
      python scripts/replay_loop.py 3 8 2 $a profile > $out/ncu_b8_$a.log 2>&1
  ncu -i $out/prof_b8_$a.ncu-rep --page raw --csv > $out/raw_b8_$a.csv 2>/dev/null
done
rm -f $out/prof_b1.ncu-rep $out/prof_b8.ncu-rep     # the raw csv pages travel back; the reports are too large
ls -la $out
