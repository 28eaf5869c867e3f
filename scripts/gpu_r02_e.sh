# round 2, call E: everything new since call D -- tests, full default bench (all legs), reference arm, sanitizers
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/e_pytest.log
timeout 900 python bench.py > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "exit $?" >> gpurun_out/e_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/e_bench_ref.json 2> gpurun_out/e_bench_ref.err; echo "exit $?" >> gpurun_out/e_bench_ref.err
timeout 400 compute-sanitizer --tool memcheck python scripts/sanitize_small.py 6 > gpurun_out/e_memcheck.log 2>&1; echo "exit $?" >> gpurun_out/e_memcheck.log
timeout 400 compute-sanitizer --tool racecheck python scripts/sanitize_small.py 6 > gpurun_out/e_racecheck.log 2>&1; echo "exit $?" >> gpurun_out/e_racecheck.log
for v in o6s16w7 o6s32w13 o5s64w13; do echo "== $v"; timeout 300 python scripts/variant_bench.py $v decompress 65536 2> gpurun_out/e_var_$v.err | python scripts/print_bench_line.py; done > gpurun_out/e_variants.log 2>&1
cat gpurun_out/e_variants.log
tail -4 gpurun_out/e_pytest.log; python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/e_bench.json") if l.startswith("{")][-1])
    print("main", d["value"], d["ms_per_step"], d["kernel_ms_per_step"], "e2e", d["e2e"], "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"])
    for k, v in (d.get("extra") or {}).items():
        print(k, v.get("value"), v.get("ms_per_step"), v.get("roofline") and (v["roofline"]["achieved"], v["roofline"]["frac"]), v.get("cpu_baseline"), v.get("ratio"), v.get("kernel_ms_per_step"))
except Exception as e:
    print("ERR", e)
try:
    d = json.loads([l for l in open("gpurun_out/e_bench_ref.json") if l.startswith("{")][-1])
    print("ref", d["value"], d["ms_per_step"], d["config"]["chunks_per_gpu"], d["cpu_baseline"]["cores"])
except Exception as e:
    print("ERR ref", e)
PY
tail -3 gpurun_out/e_bench.err; tail -2 gpurun_out/e_bench_ref.err; tail -3 gpurun_out/e_memcheck.log; tail -3 gpurun_out/e_racecheck.log
