timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s11_gpu_tests.log
for w in "c4h6 20000"; do
  set -- $w
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_kernel --launch-skip 6 -c 1 -f -o /tmp/$1_r1j python tools/profile_eloc.py $w > gpurun_out/s11_ncu_$1.log 2>&1
  ncu -i /tmp/$1_r1j.ncu-rep --page raw --csv > gpurun_out/$1_r1j_raw.csv 2>/dev/null
  ncu -i /tmp/$1_r1j.ncu-rep --page source --csv > gpurun_out/$1_r1j_source.csv 2>/dev/null
done
rm -f gpurun_out/c4h6_r1h_source.csv gpurun_out/c4h6_r1i_source.csv gpurun_out/spec_r1h.ncu-rep
tail -4 gpurun_out/s11_gpu_tests.log
