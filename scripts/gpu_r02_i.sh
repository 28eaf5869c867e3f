# round 2, call I: decode table geometry at 15 warps per SM
mkdir -p gpurun_out
for v in g6 g6b; do echo "== $v"; timeout 300 python scripts/variant_bench.py $v decompress 65536 2> gpurun_out/i_var_$v.err | python scripts/print_bench_line.py; done > gpurun_out/i_variants.log 2>&1
timeout 600 python bench.py --workload decompress --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/i_bench_dec.json 2> gpurun_out/i_bench_dec.err; echo "exit $?" >> gpurun_out/i_bench_dec.err
cat gpurun_out/i_variants.log; cat gpurun_out/i_bench_dec.json | python scripts/print_bench_line.py; tail -2 gpurun_out/i_bench_dec.err
