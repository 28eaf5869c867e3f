# round 2, call H: split-phase overflow-subtable load in the decode step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "decompress or inflate or fixture or known or gzip or reference_test" > gpurun_out/h_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/h_pytest.log
timeout 600 python bench.py --workload decompress --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/h_bench_dec.json 2> gpurun_out/h_bench_dec.err; echo "exit $?" >> gpurun_out/h_bench_dec.err
timeout 600 python bench.py --workload decompress --chunks 262144 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/h_bench_dec4.json 2> gpurun_out/h_bench_dec4.err; echo "exit $?" >> gpurun_out/h_bench_dec4.err
echo "== lf8" > gpurun_out/h_variants.log; timeout 300 python scripts/variant_bench.py lf8 decompress 65536 2> gpurun_out/h_var_lf8.err | python scripts/print_bench_line.py >> gpurun_out/h_variants.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ldb_inflate_decode' -s 1 -c 1 -o gpurun_out/prof_inflate_r02h python bench.py --workload decompress --chunks 65536 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/h_ncu_inflate.log 2>&1
tail -3 gpurun_out/h_pytest.log; cat gpurun_out/h_bench_dec.json | python scripts/print_bench_line.py; cat gpurun_out/h_bench_dec4.json | python scripts/print_bench_line.py; cat gpurun_out/h_variants.log; tail -2 gpurun_out/h_bench_dec.err
