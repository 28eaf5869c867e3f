mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/bench_rt.json 2> gpurun_out/bench_rt.err; echo "exit $?" >> gpurun_out/bench_rt.err
timeout 600 python bench.py --workload decompress --steps 6 --warmup 3 --no-cpu > gpurun_out/bench_dec.json 2> gpurun_out/bench_dec.err; echo "exit $?" >> gpurun_out/bench_dec.err
for s in 2 8; do LIBDEFLATE_B200_PIPE_STAGES=$s timeout 600 python bench.py --workload decompress --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_dec_s$s.json 2> gpurun_out/bench_dec_s$s.err; done
tail -3 gpurun_out/pytest_gpu.log
for f in bench_rt bench_dec bench_dec_s2 bench_dec_s8; do echo "== $f"; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/$f.json"))
    print({k: d.get(k) for k in ("value", "ms_per_step", "kernel_ms_per_step")}, d.get("e2e"))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/$f.err").read()[-1500:])
PY
done
