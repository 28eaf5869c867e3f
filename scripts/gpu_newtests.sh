mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "bgzf" > gpurun_out/pytest_bgzf.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_bgzf.log
tail -4 gpurun_out/pytest_bgzf.log
