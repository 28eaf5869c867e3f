# round 2, call K: decode kernel with the code-length scratch in global memory; warps-per-SM variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "decompress or inflate or fixture or known or gzip or reference_test" > gpurun_out/k_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/k_pytest.log
timeout 600 python bench.py --workload decompress --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/k_bench_dec.json 2> gpurun_out/k_bench_dec.err; echo "exit $?" >> gpurun_out/k_bench_dec.err
for v in w18 w21 w24; do echo "== $v"; timeout 300 python scripts/variant_bench.py $v decompress 65536 2> gpurun_out/k_var_$v.err | python scripts/print_bench_line.py; timeout 300 python scripts/variant_bench.py $v decompress 131072 2>> gpurun_out/k_var_$v.err | python scripts/print_bench_line.py; done > gpurun_out/k_variants.log 2>&1
timeout 300 python bench.py --workload decompress --chunks 131072 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/k_bench_dec2.json 2> gpurun_out/k_bench_dec2.err
tail -3 gpurun_out/k_pytest.log; cat gpurun_out/k_bench_dec.json | python scripts/print_bench_line.py; cat gpurun_out/k_bench_dec2.json | python scripts/print_bench_line.py; cat gpurun_out/k_variants.log; tail -2 gpurun_out/k_bench_dec.err
