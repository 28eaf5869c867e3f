mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload decompress --steps 10 --warmup 3 > gpurun_out/bench_dec.json 2> gpurun_out/bench_dec.err; echo "exit $?" >> gpurun_out/bench_dec.err
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_rt.json 2> gpurun_out/bench_rt.err; echo "exit $?" >> gpurun_out/bench_rt.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "exit $?" >> gpurun_out/bench_ref.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ldb_inflate_kernel -s 3 -c 1 -o gpurun_out/prof_inflate21 python bench.py --workload decompress --chunks 32768 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_inflate.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; for f in bench_dec bench_rt bench_ref; do echo "== $f"; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/$f.json"))
    print({k: d.get(k) for k in ("value", "ms_per_step", "ratio", "kernel_ms_per_step", "roofline")}, d.get("e2e") and d["e2e"]["value"], d.get("cpu_baseline") and {k: d["cpu_baseline"].get(k) for k in ("value","cores","cpu_quota","compress_MBps","decompress_MBps","ratio")})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/$f.err").read()[-1500:])
PY
done
