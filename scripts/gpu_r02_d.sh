# round 2, call D: unified decode step + 32 resolve warps per SM; full default bench for reference
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/d_pytest.log
timeout 600 python bench.py --workload decompress --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/d_bench_dec.json 2> gpurun_out/d_bench_dec.err; echo "exit $?" >> gpurun_out/d_bench_dec.err
for v in q128 q1024; do
  echo "== $v"
  timeout 300 python scripts/variant_bench.py $v decompress 65536 2> gpurun_out/d_var_$v.err | python -c "
import sys, json
for line in sys.stdin:
    line = line.rstrip()
    if line.startswith('{'):
        d = json.loads(line); print(d['value'], d['kernel_ms_per_step'])
"
done > gpurun_out/d_variants.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ldb_inflate_(decode|resolve)' -s 2 -c 2 -o gpurun_out/prof_inflate_r02d python bench.py --workload decompress --chunks 65536 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/d_ncu_inflate.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/d_bench_rt.json 2> gpurun_out/d_bench_rt.err; echo "exit $?" >> gpurun_out/d_bench_rt.err
tail -3 gpurun_out/d_pytest.log; python - <<'PY'
import json
for f in ("d_bench_dec", "d_bench_rt"):
    try:
        d = json.loads([l for l in open("gpurun_out/%s.json" % f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d["kernel_ms_per_step"], d["roofline"]["frac"], "e2e", d.get("e2e") and d["e2e"]["value"], "cpu", d.get("cpu_baseline") and d["cpu_baseline"]["value"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/d_bench_dec.err; cat gpurun_out/d_variants.log; tail -3 gpurun_out/d_bench_rt.err
