# round 2, call M: cache-streaming hints for the token stream -- time and DRAM traffic, default vs no hints vs 24 resolve warps
mkdir -p gpurun_out
K='ldb_(inflate_decode_kernel|inflate_resolve_kernel)'
timeout 600 python bench.py --workload decompress --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/m_bench_dec.json 2> gpurun_out/m_bench_dec.err
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:"$K" -s 2 -c 2 --csv --log-file gpurun_out/m_traffic_default.csv python bench.py --workload decompress --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/m_t0.log 2>&1
for v in nohint r24; do
  echo "== $v"; timeout 300 python scripts/variant_bench.py $v decompress 65536 2> gpurun_out/m_var_$v.err | python scripts/print_bench_line.py
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:"$K" -s 2 -c 2 --csv --log-file gpurun_out/m_traffic_$v.csv python scripts/variant_bench.py $v decompress 65536 > gpurun_out/m_t_$v.log 2>&1
done > gpurun_out/m_variants.log 2>&1
cat gpurun_out/m_bench_dec.json | python scripts/print_bench_line.py; cat gpurun_out/m_variants.log
for f in default nohint r24; do echo "-- $f"; python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/m_traffic_$f.csv", errors="replace")) if len(r)>5]
hdr=next(r for r in rows if "Kernel Name" in r)
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rows:
    if r is hdr or r[ik]=="Kernel Name": continue
    print(r[ik].split("(")[0], r[im], r[iv], r[iu])
PY
done
