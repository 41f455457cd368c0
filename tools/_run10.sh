timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s13_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s13_smoke.log 2>&1
( for w in "c4h6 20000" "h2o 100000" "lih 1000000"; do timeout 300 python tools/time_kernels.py $w 2>&1 | tail -1; done
timeout 300 python tools/gpu_config4.py 2>&1 | tail -3 ) > gpurun_out/s13_time.log 2>&1
timeout 300 python bench.py > gpurun_out/s13_bench_n1.json 2> gpurun_out/s13_bench_n1.err
tail -3 gpurun_out/s13_gpu_tests.log; tail -3 gpurun_out/s13_smoke.log; cat gpurun_out/s13_time.log; cat gpurun_out/s13_bench_n1.json
