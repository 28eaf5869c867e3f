# round 2, call R: phase clocks of the deflate kernel (variant build with -DLZ_TIMING), 16384 x 64 KiB, L6
mkdir -p gpurun_out
timeout 600 python scripts/variant_bench.py timing roundtrip 16384 > gpurun_out/r_timing.log 2> gpurun_out/r_timing.err; echo "exit $?" >> gpurun_out/r_timing.err
grep timing gpurun_out/r_timing.log; tail -2 gpurun_out/r_timing.err
