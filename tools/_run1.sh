timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s3_gpu_tests.log
for w in "lih 1000000" "h2 1000000" "lih_sto 1000000" "h2o 100000" "c4h6 20000"; do timeout 300 python tools/time_kernels.py $w 2>&1 | tail -2; done > gpurun_out/s3_time.log
QMCB_SPEC_DEFS="-DSPEC_MINB_ELOC=4" timeout 300 python tools/time_kernels.py lih 1000000 2>&1 | tail -1 >> gpurun_out/s3_time.log
QMCB_MB8=1 timeout 300 python tools/time_kernels.py c4h6 20000 2>&1 | tail -1 >> gpurun_out/s3_time.log
tail -5 gpurun_out/s3_gpu_tests.log; cat gpurun_out/s3_time.log
