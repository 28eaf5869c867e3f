# round 2, call B: slim decode loop + warp-per-chunk resolve
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b_pytest.log
timeout 600 python bench.py --workload decompress --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_bench_dec.json 2> gpurun_out/b_bench_dec.err; echo "exit $?" >> gpurun_out/b_bench_dec.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ldb_inflate_(decode|resolve)' -s 2 -c 2 -o gpurun_out/prof_inflate_r02b python bench.py --workload decompress --chunks 65536 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_ncu_inflate.log 2>&1
timeout 400 compute-sanitizer --tool racecheck python scripts/sanitize_small.py 6 > gpurun_out/b_racecheck.log 2>&1; echo "exit $?" >> gpurun_out/b_racecheck.log
tail -3 gpurun_out/b_pytest.log; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/b_bench_dec.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["kernel_ms_per_step"], d["roofline"]["frac"])
PY
tail -3 gpurun_out/b_bench_dec.err; tail -4 gpurun_out/b_racecheck.log
