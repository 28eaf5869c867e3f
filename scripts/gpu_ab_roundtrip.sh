mkdir -p gpurun_out
for v in q0 q6 q8 q12 q0; do
  echo "== $v"
  timeout 300 python scripts/variant_bench.py $v roundtrip 16384 2> gpurun_out/var_$v.err | python -c "
import sys, json
for line in sys.stdin:
    line = line.rstrip()
    if line.startswith('{'):
        d = json.loads(line); print(d['value'], d['kernel_ms_per_step'], d.get('verified'))
    else:
        print(line)
"
done
