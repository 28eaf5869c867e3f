# round 2, call Y: resolve kernel with un-wrapped staging stores (ring slack); reference point 8.97 ms
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "decompress or inflate or fixture or known or gzip or reference_test or large" > gpurun_out/y_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/y_pytest.log
timeout 300 python bench.py --workload decompress --steps 5 --warmup 3 --no-e2e --no-cpu 2> gpurun_out/y_default.err | python scripts/print_bench_line.py > gpurun_out/y_bench.log
tail -3 gpurun_out/y_pytest.log; cat gpurun_out/y_bench.log
