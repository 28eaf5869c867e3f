mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_n2b.json 2> gpurun_out/bench_n2b.err; echo "exit $?" >> gpurun_out/bench_n2b.err
echo "stdout lines: $(wc -l < gpurun_out/bench_n2b.json)"; head -c 200 gpurun_out/bench_n2b.json; echo; tail -3 gpurun_out/bench_n2b.err | cut -c1-200
