# round 2, call T: decode takes a literal that follows a literal OR a completed match in the same step; variant with two follow-on literals
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "decompress or inflate or fixture or known or gzip or reference_test" > gpurun_out/t_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/t_pytest.log
timeout 600 python bench.py --workload decompress --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/t_bench_dec.json 2> gpurun_out/t_bench_dec.err; echo "exit $?" >> gpurun_out/t_bench_dec.err
for v in lit3; do echo "== $v"; timeout 300 python scripts/variant_bench.py $v decompress 65536 2> gpurun_out/t_var_$v.err | python scripts/print_bench_line.py; done > gpurun_out/t_variants.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ldb_inflate_decode' -s 1 -c 1 -o gpurun_out/prof_inflate_r02t python bench.py --workload decompress --chunks 65536 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/t_ncu_inflate.log 2>&1
tail -3 gpurun_out/t_pytest.log; cat gpurun_out/t_bench_dec.json | python scripts/print_bench_line.py; cat gpurun_out/t_variants.log; tail -2 gpurun_out/t_bench_dec.err
