mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1; lscpu > gpurun_out/lscpu.txt 2>&1; free -g >> gpurun_out/lscpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload decompress --steps 5 --warmup 3 > gpurun_out/bench_dec.json 2> gpurun_out/bench_dec.err; echo "exit $?" >> gpurun_out/bench_dec.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ldb_inflate -s 3 -c 1 -o gpurun_out/prof_inflate python bench.py --workload decompress --chunks 16384 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_inflate.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_dec.json | cut -c1-3000; tail -3 gpurun_out/bench_dec.err
