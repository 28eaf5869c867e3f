mkdir -p gpurun_out
for v in inf_w15 inf_w13 inf_w12 inf_w11 inf_w9; do
  echo "== $v"
  timeout 300 python scripts/variant_bench.py $v decompress 65536 2> gpurun_out/var_$v.err | python -c "
import sys, json
for line in sys.stdin:
    line = line.rstrip()
    if line.startswith('{'):
        d = json.loads(line); print(d['value'], d['kernel_ms_per_step'], d.get('verified'))
    else:
        print(line)
"
done
