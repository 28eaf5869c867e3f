# round 2, call U: how many follow-on literals per decode step, with and without the two-slot lookahead
mkdir -p gpurun_out
for v in lit3 lit4 lit5 lit3w lit4w; do echo "== $v"; timeout 300 python scripts/variant_bench.py $v decompress 65536 2> gpurun_out/u_var_$v.err | python scripts/print_bench_line.py; done > gpurun_out/u_variants.log 2>&1
cat gpurun_out/u_variants.log
