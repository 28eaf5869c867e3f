mkdir -p gpurun_out
for cfg in "12 1048576 512" "9 65536 8192" "1 65536 16384"; do
  set -- $cfg
  echo "== level $1 chunk $2 n $3"
  timeout 600 python scripts/variant_bench.py tim roundtrip $3 $1 $2 2> gpurun_out/var_tim_$1.err | python -c "
import sys, json
for line in sys.stdin:
    line = line.rstrip()
    if line.startswith('{'):
        d = json.loads(line); print(d['value'], d['kernel_ms_per_step'], d.get('ratio'))
    else:
        print(line)
"
done
for v in prev exsm prev exsm; do
  echo "== $v"
  timeout 300 python scripts/variant_bench.py $v roundtrip 16384 2> gpurun_out/var_$v.err | python -c "
import sys, json
for line in sys.stdin:
    line = line.rstrip()
    if line.startswith('{'):
        d = json.loads(line); print(d['value'], d['kernel_ms_per_step'])
    else:
        print(line)
"
done
