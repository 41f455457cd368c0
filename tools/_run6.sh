timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s8_gpu_tests.log
( for w in "c4h6 20000" "h2o 100000" "lih 1000000" "h2 1000000" "lih_sto 1000000"; do timeout 300 python tools/time_kernels.py $w 2>&1 | tail -1; done
timeout 300 python tools/gpu_config4.py 2>&1 | tail -3 ) > gpurun_out/s8_time.log 2>&1
timeout 300 python bench.py > gpurun_out/s8_bench_n1.json 2> gpurun_out/s8_bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s8_bench_reference.json 2>> gpurun_out/s8_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s8_launches_full.csv python bench.py --steps 5 --warmup 3 --therm 3 --no-cpu-baseline > gpurun_out/s8_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spec_ --launch-skip 6 -c 2 -f -o gpurun_out/spec_r1i python tools/profile_eloc.py lih 1000000 > gpurun_out/s8_ncu.log 2>&1
ncu -i gpurun_out/spec_r1i.ncu-rep --page raw --csv > gpurun_out/spec_r1i_raw.csv 2>/dev/null
for w in "c4h6 20000"; do
  set -- $w
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_kernel --launch-skip 6 -c 1 -f -o /tmp/$1_r1i python tools/profile_eloc.py $w > gpurun_out/s8_ncu_$1.log 2>&1
  ncu -i /tmp/$1_r1i.ncu-rep --page raw --csv > gpurun_out/$1_r1i_raw.csv 2>/dev/null
  ncu -i /tmp/$1_r1i.ncu-rep --page source --csv > gpurun_out/$1_r1i_source.csv 2>/dev/null
done
rm -f gpurun_out/*_r1h_source.csv
tail -3 gpurun_out/s8_gpu_tests.log; cat gpurun_out/s8_time.log; cat gpurun_out/s8_bench_n1.json; ls -la gpurun_out
