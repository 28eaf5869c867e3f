mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "random_mix or stored_blocks or native_library" > gpurun_out/pytest_new.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_new.log
tail -4 gpurun_out/pytest_new.log
