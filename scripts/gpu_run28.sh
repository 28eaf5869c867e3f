mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "exit $?" >> gpurun_out/bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "exit $?" >> gpurun_out/bench_ref_n2.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cat gpurun_out/gpus.txt
for f in bench_n2 bench_ref_n2 bench_n1; do echo "== $f"; python - <<PY
import json
try:
    txt = [l for l in open("gpurun_out/$f.json") if l.startswith("{")][-1]
    d = json.loads(txt)
    print({k: d.get(k) for k in ("impl", "value", "n_gpus", "ms_per_step", "scaling", "kernel_ms_per_step")}, d.get("e2e"))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/$f.err").read()[-2500:])
PY
done
