# 2 GPUs of one box: the NCCL data-plane test, two devices in one process, bench at N=2 (pre-sharded + single origin)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/n2_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "two_gpu or two_devices or two_contexts" > gpurun_out/n2_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/n2_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo "exit $?" >> gpurun_out/n2_bench.err
tail -4 gpurun_out/n2_pytest.log
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/n2_bench.json") if l.startswith("{")][-1])
    print("N=2", d["value"], d["ms_per_step"], "e2e", d["e2e"] and d["e2e"]["value"], "single_origin", d.get("single_origin"))
except Exception as e:
    print("ERR", e)
PY
tail -5 gpurun_out/n2_bench.err
