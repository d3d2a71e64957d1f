export REGTR_GEMM_WIDE=1
mkdir -p gpurun_out; : > gpurun_out/ab.txt
for e in "A=1" "REGTR_GEMM_LOWSPLIT=1"; do
  env $e timeout 250 python bench.py --steps 100 --cpu-baseline 0 --checks 1 > /tmp/o.json 2> /tmp/o.err
  python - "$e" >> gpurun_out/ab.txt <<'PY'
import sys, json
try:
    d = json.loads(open('/tmp/o.json').read())
    print(sys.argv[1], round(d['value'], 1), round(d['e2e']['value'], 1), d['pose_err_vs_oracle'])
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex); print(open('/tmp/o.err').read()[-300:])
PY
done
cat gpurun_out/ab.txt
