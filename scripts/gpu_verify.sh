mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_rt.json 2> gpurun_out/bench_rt.err; echo "exit $?" >> gpurun_out/bench_rt.err
tail -2 gpurun_out/smoke.log; tail -3 gpurun_out/pytest_gpu.log; python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_rt.json") if l.startswith("{")][-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "ratio", "kernel_ms_per_step")}, "e2e", d["e2e"]["value"], "launches", d.get("gpu_launches"))
PY
