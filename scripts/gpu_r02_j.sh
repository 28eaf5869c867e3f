# round 2, call J: deflate tuning variants (time and ratio), 16384 chunks
mkdir -p gpurun_out
for v in t512 run8 run8d32 d16; do echo "== $v"; timeout 300 python scripts/variant_bench.py $v roundtrip 16384 2> gpurun_out/j_var_$v.err | python -c "
import sys, json
for line in sys.stdin:
    line = line.rstrip()
    if line.startswith('{'):
        d = json.loads(line); print(d['value'], d['kernel_ms_per_step'], 'ratio', d['ratio'])
    elif line: print(line)
"; done > gpurun_out/j_variants.log 2>&1
timeout 300 python bench.py --chunks 16384 --steps 3 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/j_base.json 2> gpurun_out/j_base.err
cat gpurun_out/j_variants.log; python -c "
import json
d = json.loads([l for l in open('gpurun_out/j_base.json') if l.startswith('{')][-1]); print('base', d['value'], d['kernel_ms_per_step'], 'ratio', d['ratio'])"
tail -3 gpurun_out/j_var_t512.err
