# round 2, call A: first GPU contact of the two-kernel inflate + the LZ_QUANTUM A/B left over from round 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/a_smoke.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/a_pytest.log
timeout 600 python bench.py --workload decompress --steps 5 --warmup 3 --no-e2e > gpurun_out/a_bench_dec.json 2> gpurun_out/a_bench_dec.err; echo "exit $?" >> gpurun_out/a_bench_dec.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ldb_inflate_(decode|resolve)' -s 2 -c 2 -o gpurun_out/prof_inflate_r02a python bench.py --workload decompress --chunks 16384 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/a_ncu_inflate.log 2>&1
bash scripts/gpu_ab_roundtrip.sh > gpurun_out/a_ab_quantum.log 2>&1
timeout 400 compute-sanitizer --tool memcheck python scripts/sanitize_small.py 6 > gpurun_out/a_memcheck.log 2>&1; echo "exit $?" >> gpurun_out/a_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python scripts/sanitize_small.py 6 > gpurun_out/a_racecheck.log 2>&1; echo "exit $?" >> gpurun_out/a_racecheck.log
tail -3 gpurun_out/a_smoke.log; tail -3 gpurun_out/a_pytest.log; cat gpurun_out/a_bench_dec.json | cut -c1-1500; tail -3 gpurun_out/a_bench_dec.err; cat gpurun_out/a_ab_quantum.log; tail -5 gpurun_out/a_memcheck.log; tail -5 gpurun_out/a_racecheck.log
