# round 2, call C: resolve with the window in L2 (16 warps/SM) + occupancy variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c_pytest.log
timeout 600 python bench.py --workload decompress --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/c_bench_dec.json 2> gpurun_out/c_bench_dec.err; echo "exit $?" >> gpurun_out/c_bench_dec.err
for v in r8 r24 r32 l3 l6; do
  echo "== $v"
  timeout 300 python scripts/variant_bench.py $v decompress 65536 2> gpurun_out/c_var_$v.err | python -c "
import sys, json
for line in sys.stdin:
    line = line.rstrip()
    if line.startswith('{'):
        d = json.loads(line); print(d['value'], d['kernel_ms_per_step'])
"
done > gpurun_out/c_variants.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ldb_inflate_resolve' -s 1 -c 1 -o gpurun_out/prof_inflate_r02c python bench.py --workload decompress --chunks 65536 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/c_ncu_inflate.log 2>&1
tail -3 gpurun_out/c_pytest.log; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c_bench_dec.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["kernel_ms_per_step"], d["roofline"]["frac"])
PY
tail -3 gpurun_out/c_bench_dec.err; cat gpurun_out/c_variants.log
