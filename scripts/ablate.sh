# marginal cost of each stage under 6-way overlap (diagnosis; results are numerically meaningless)
mkdir -p gpurun_out; : > gpurun_out/ablate.txt
for a in none mha agg gemm norm ln bq_up "mha,agg,gemm,norm,ln,bq_up"; do
  v=$(REGTR_ABLATE=$a timeout 200 python bench.py --steps 60 --warmup 5 --cpu-baseline 0 --checks 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(1000/d['value'],3), 'ms/pair')")
  echo "ablate=$a -> $v" >> gpurun_out/ablate.txt
done
cat gpurun_out/ablate.txt
