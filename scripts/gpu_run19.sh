mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_rt.json 2> gpurun_out/bench_rt.err; echo "exit $?" >> gpurun_out/bench_rt.err
timeout 600 python bench.py --level 1 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_l1.json 2> gpurun_out/bench_l1.err
timeout 600 python bench.py --level 9 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_l9.json 2> gpurun_out/bench_l9.err
timeout 900 python bench.py --level 12 --chunks 4096 --chunk-size 1048576 --steps 2 --warmup 3 --no-e2e > gpurun_out/bench_l12.json 2> gpurun_out/bench_l12.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ldb_deflate_lz -s 3 -c 1 -o gpurun_out/prof_deflate19 python bench.py --chunks 4096 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_deflate.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; for f in bench_rt bench_l1 bench_l9 bench_l12; do echo "== $f"; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/$f.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "ratio", "kernel_ms_per_step")}, d.get("e2e") and d["e2e"]["value"], d.get("cpu_baseline") and {k: d["cpu_baseline"].get(k) for k in ("value","cores","cpu_quota","compress_MBps","decompress_MBps","ratio")})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/$f.err").read()[-1500:])
PY
done
