timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s9_gpu_tests.log
( for w in "c4h6 20000" "h2o 100000" "lih 1000000"; do timeout 300 python tools/time_kernels.py $w 2>&1 | tail -1; done
QMCB_CTA_THREADS=128 timeout 300 python tools/time_kernels.py c4h6 20000 2>&1 | tail -1
QMCB_MB8=1 timeout 300 python tools/time_kernels.py c4h6 20000 2>&1 | tail -1
timeout 300 python tools/gpu_config4.py 2>&1 | tail -3 ) > gpurun_out/s9_time.log 2>&1
tail -3 gpurun_out/s9_gpu_tests.log; cat gpurun_out/s9_time.log
