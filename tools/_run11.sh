timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s15_gpu_tests.log
( for w in "c4h6 20000" "h2o 100000" "lih 1000000"; do timeout 300 python tools/time_kernels.py $w 2>&1 | tail -1; done ) > gpurun_out/s15_time.log 2>&1
timeout 300 python tools/gpu_check.py 2>&1 | grep -E "c4h6|h2o_ground" > gpurun_out/s15_check.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/s15_bench_n1.json 2> gpurun_out/s15_bench_n1.err
tail -4 gpurun_out/s15_gpu_tests.log; cat gpurun_out/s15_time.log gpurun_out/s15_check.log; cut -c1-120 gpurun_out/s15_bench_n1.json; python -c "
import json; d=json.load(open('gpurun_out/s15_bench_n1.json')); print(d['e2e'], d['roofline']['frac'])"
