timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s12_gpu_tests.log
( timeout 300 python tools/gpu_config4.py 2>&1 | tail -3
for w in "c4h6 20000" "h2o 100000" "lih 1000000"; do timeout 300 python tools/time_kernels.py $w 2>&1 | tail -1; done ) > gpurun_out/s12_time.log 2>&1
timeout 300 python tools/gpu_check.py 2>&1 | grep -E "een|c4h6" > gpurun_out/s12_check.log
tail -4 gpurun_out/s12_gpu_tests.log; cat gpurun_out/s12_time.log gpurun_out/s12_check.log
