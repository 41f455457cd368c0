timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s4_gpu_tests.log
timeout 300 python tools/time_kernels.py lih 1000000 2>&1 | tail -1 > gpurun_out/s4_time.log
QMCB_LIB=$PWD/qmctorch_b200/lib/libqmcb_e10.so QMCB_SPEC_DEFS="-DQMCB_ETAB_LOG2=10" timeout 300 python tools/time_kernels.py lih 1000000 2>&1 | tail -1 >> gpurun_out/s4_time.log
QMCB_LIB=$PWD/qmctorch_b200/lib/libqmcb_e10.so timeout 300 python tools/time_kernels.py c4h6 20000 2>&1 | tail -1 >> gpurun_out/s4_time.log
QMCB_SPEC_DEFS="-DSPEC_MINB_ELOC=4" timeout 300 python tools/time_kernels.py lih 1000000 2>&1 | tail -1 >> gpurun_out/s4_time.log
QMCB_SPEC_DEFS="-DSPEC_MINB_ELOC=2" timeout 300 python tools/time_kernels.py lih 1000000 2>&1 | tail -1 >> gpurun_out/s4_time.log
QMCB_SPEC_THREADS=64 QMCB_SPEC_DEFS="-DSPEC_MINB_ELOC=6" timeout 300 python tools/time_kernels.py lih 1000000 2>&1 | tail -1 >> gpurun_out/s4_time.log
timeout 300 python bench.py > gpurun_out/s4_bench_n1.json 2> gpurun_out/s4_bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spec_eloc -c 1 -f -o gpurun_out/spec_r1h python tools/profile_eloc.py lih 1000000 > gpurun_out/s4_ncu.log 2>&1
tail -3 gpurun_out/s4_gpu_tests.log; cat gpurun_out/s4_time.log; cat gpurun_out/s4_bench_n1.json
