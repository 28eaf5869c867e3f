# End-of-round verification on one B200: smoke, GPU parity tests, the bench lines of every BASELINE config that fits
# one GPU, the reference arm, DRAM traffic (ncu) for the roofline.traffic field, and one full ncu capture.
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'ldb_(inflate_kernel|deflate_lz_kernel)' -c 4 --csv --log-file gpurun_out/traffic_rt.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/traffic_rt.log 2>&1
python scripts/ncu_traffic.py gpurun_out/traffic_rt.csv roundtrip_L6_65536x65536 > gpurun_out/traffic_rt.json 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'ldb_(inflate_kernel|deflate_lz_kernel)' -c 2 --csv --log-file gpurun_out/traffic_dec.csv python bench.py --workload decompress --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/traffic_dec.log 2>&1
python scripts/ncu_traffic.py gpurun_out/traffic_dec.csv decompress_L6_65536x65536 > gpurun_out/traffic_dec.json 2>&1
cp profiles/dram_traffic.json gpurun_out/dram_traffic.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_rt.json 2> gpurun_out/bench_rt.err; echo "exit $?" >> gpurun_out/bench_rt.err
timeout 600 python bench.py --workload decompress --steps 10 --warmup 3 > gpurun_out/bench_dec.json 2> gpurun_out/bench_dec.err; echo "exit $?" >> gpurun_out/bench_dec.err
timeout 600 python bench.py --level 1 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_l1.json 2> gpurun_out/bench_l1.err
timeout 600 python bench.py --level 9 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_l9.json 2> gpurun_out/bench_l9.err
timeout 900 python bench.py --level 12 --chunks 4096 --chunk-size 1048576 --steps 2 --warmup 3 --no-e2e > gpurun_out/bench_l12.json 2> gpurun_out/bench_l12.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ldb_deflate_lz -s 3 -c 1 -o gpurun_out/prof_deflate_final python bench.py --chunks 4096 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_deflate.log 2>&1
cat gpurun_out/smoke.log | tail -2; tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/traffic_rt.json gpurun_out/traffic_dec.json
for f in bench_ref bench_rt bench_dec bench_l1 bench_l9 bench_l12; do echo "== $f"; python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/$f.json") if l.startswith("{")][-1])
    print({k: d.get(k) for k in ("impl", "value", "ms_per_step", "ratio", "kernel_ms_per_step")}, "roofline", d.get("roofline") and {k: d["roofline"].get(k) for k in ("kernel","achieved","frac","traffic")}, "e2e", d.get("e2e") and d["e2e"]["value"], "cpu", d.get("cpu_baseline") and {k: d["cpu_baseline"].get(k) for k in ("value","cores","compress_MBps","decompress_MBps","ratio")}, "clocks", d.get("clocks"))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/$f.err").read()[-1500:])
PY
done
