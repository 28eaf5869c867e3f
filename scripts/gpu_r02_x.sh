# round 2, call X: compute-sanitizer over the final kernels (the decode step changed after the first sanitizer run)
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck python scripts/sanitize_small.py 6 > gpurun_out/x_memcheck.log 2>&1; echo "exit $?" >> gpurun_out/x_memcheck.log
timeout 500 compute-sanitizer --tool racecheck python scripts/sanitize_small.py 6 > gpurun_out/x_racecheck.log 2>&1; echo "exit $?" >> gpurun_out/x_racecheck.log
tail -4 gpurun_out/x_memcheck.log; tail -4 gpurun_out/x_racecheck.log
