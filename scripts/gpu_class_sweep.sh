# robustness sweep: the round trip and the decompress-only workload on every synthetic chunk class (16384 x 64 KiB)
mkdir -p gpurun_out
: > gpurun_out/sweep.log
for c in 0 1 2 3 4 5 6; do
  echo "== class $c roundtrip" >> gpurun_out/sweep.log
  timeout 300 python bench.py --data-class $c --chunks 16384 --steps 3 --warmup 3 --no-e2e --no-cpu --no-extra 2>> gpurun_out/sweep.err | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line); print(d['value'], d['kernel_ms_per_step'], 'ratio', d['ratio'])
" >> gpurun_out/sweep.log
  echo "== class $c decompress (reference L6 streams)" >> gpurun_out/sweep.log
  timeout 300 python bench.py --workload decompress --data-class $c --chunks 16384 --steps 3 --warmup 3 --no-e2e --no-cpu 2>> gpurun_out/sweep.err | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line); print(d['value'], d['kernel_ms_per_step'], 'ratio_ref', d['ratio_reference_L6'])
" >> gpurun_out/sweep.log
done
cat gpurun_out/sweep.log; tail -3 gpurun_out/sweep.err
