# round 2, call W: decode variants on top of the multi-symbol step (default: 4 follow-on literals, 5-bit main offset table)
mkdir -p gpurun_out
timeout 300 python bench.py --workload decompress --steps 5 --warmup 3 --no-e2e --no-cpu 2> gpurun_out/w_default.err | python scripts/print_bench_line.py > gpurun_out/w_variants.log
for v in lit6 g6 g6lit6; do echo "== $v"; timeout 300 python scripts/variant_bench.py $v decompress 65536 2> gpurun_out/w_var_$v.err | python scripts/print_bench_line.py; done >> gpurun_out/w_variants.log 2>&1
cat gpurun_out/w_variants.log
