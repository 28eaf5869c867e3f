# round 2, call Q: e2e with the two-host-thread form (compress of group g+1 over decompress of group g)
mkdir -p gpurun_out
LDB_E2E_TRACE=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-extra --no-l12 > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "exit $?" >> gpurun_out/q_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/q_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["kernel_ms_per_step"], d["e2e"])
PY
grep -i "timeline" gpurun_out/q_bench.err; tail -2 gpurun_out/q_bench.err
