mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "gpus $N; mem: $(free -g | awk '/Mem/{print $2" GiB total, "$7" GiB avail"}'); cpu.max $(cat /sys/fs/cgroup/cpu.max 2>/dev/null)" > gpurun_out/scale_env.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_nN.json 2> gpurun_out/bench_nN.err; echo "exit $?" >> gpurun_out/bench_nN.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_nN.json 2> gpurun_out/bench_ref_nN.err; echo "exit $?" >> gpurun_out/bench_ref_nN.err
cat gpurun_out/scale_env.txt
for f in bench_nN bench_ref_nN; do echo "== $f"; python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/$f.json") if l.startswith("{")][-1])
    print({k: d.get(k) for k in ("impl", "value", "n_gpus", "ms_per_step", "scaling", "kernel_ms_per_step")}, d.get("e2e"))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/$f.err").read()[-2500:])
PY
done
