mkdir -p gpurun_out
for v in i_base i_r3 i_r6 i_m4 i_m16 i_c8 i_r6m16 i_base; do
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
