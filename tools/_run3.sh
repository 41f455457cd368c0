timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s5_gpu_tests.log
T="timeout 300 python tools/time_kernels.py lih 1000000"
( $T | tail -1
QMCB_SPEC_DEFS="-DSPEC_EUNROLL=2" $T | tail -1
QMCB_SPEC_DEFS="-DSPEC_EUNROLL=2 -DSPEC_MINB_ELOC=2" $T | tail -1
QMCB_SPEC_DEFS="-DSPEC_EUNROLL=4 -DSPEC_MINB_ELOC=2" $T | tail -1
QMCB_SPEC_DEFS="-DSPEC_PREFETCH=1" $T | tail -1
QMCB_SPEC_THREADS=96 $T | tail -1
QMCB_SPEC_THREADS=256 QMCB_SPEC_DEFS="-DSPEC_MINB_ELOC=1" $T | tail -1
) > gpurun_out/s5_time.log 2>&1
timeout 300 python bench.py > gpurun_out/s5_bench_n1.json 2> gpurun_out/s5_bench_n1.err
QMCB_STATS_2STAGE=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/s5_bench_2stage.json 2>> gpurun_out/s5_bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_kernel --launch-skip 6 -c 2 -f -o gpurun_out/c4h6_r1h python tools/profile_eloc.py c4h6 20000 > gpurun_out/s5_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_kernel --launch-skip 6 -c 2 -f -o gpurun_out/h2o_r1h python tools/profile_eloc.py h2o 100000 >> gpurun_out/s5_ncu.log 2>&1
tail -3 gpurun_out/s5_gpu_tests.log; cat gpurun_out/s5_time.log; cat gpurun_out/s5_bench_n1.json gpurun_out/s5_bench_2stage.json
