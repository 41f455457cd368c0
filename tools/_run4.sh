timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s6_gpu_tests.log
T="timeout 300 python tools/time_kernels.py lih 1000000"
( $T | tail -1
QMCB_SPEC_DEFS="-DSPEC_MINB_ELOC=4" $T | tail -1
QMCB_SPEC_THREADS=64 QMCB_SPEC_DEFS="-DSPEC_MINB_ELOC=6" $T | tail -1
QMCB_SPEC_THREADS=96 QMCB_SPEC_DEFS="-DSPEC_MINB_ELOC=4" $T | tail -1
timeout 300 python tools/time_kernels.py h2 1000000 | tail -1
timeout 300 python tools/time_kernels.py lih_sto 1000000 | tail -1
) > gpurun_out/s6_time.log 2>&1
timeout 300 python bench.py > gpurun_out/s6_bench_n1.json 2> gpurun_out/s6_bench_n1.err
for w in "c4h6 20000" "h2o 100000"; do
  set -- $w
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_kernel --launch-skip 6 -c 1 -f -o /tmp/$1_r1h python tools/profile_eloc.py $w > gpurun_out/s6_ncu_$1.log 2>&1
  ncu -i /tmp/$1_r1h.ncu-rep --page raw --csv > gpurun_out/$1_r1h_raw.csv 2>/dev/null
  ncu -i /tmp/$1_r1h.ncu-rep --page source --csv > gpurun_out/$1_r1h_source.csv 2>/dev/null
  ls -la /tmp/$1_r1h.ncu-rep >> gpurun_out/s6_ncu_$1.log
done
tail -3 gpurun_out/s6_gpu_tests.log; cat gpurun_out/s6_time.log; cat gpurun_out/s6_bench_n1.json; ls -la gpurun_out
