# 8 GPUs of one box: bench at N=8 (pre-sharded value + e2e + single-origin NCCL scatter/gather of 524288 x 64 KiB)
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/n${N}_smi.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err; echo "exit $?" >> gpurun_out/n${N}_bench.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/n${N}_bench.json") if l.startswith("{")][-1])
    print("N=${N}", d["value"], d["ms_per_step"], "e2e", d["e2e"] and d["e2e"]["value"], "single_origin", d.get("single_origin"))
except Exception as e:
    print("ERR", e)
PY
tail -8 gpurun_out/n${N}_bench.err
