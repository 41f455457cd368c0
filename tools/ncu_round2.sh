#!/bin/bash
# Round-2 profiler captures (run on a B200 box from the repository root): launch list of one bench run and
# `ncu --set full` captures of the dominant kernels, exported as CSV into gpurun_out/ (summaries are copied
# to profiles/ by hand).  Numbers printed by runs under ncu are never bench values.
out=gpurun_out
mkdir -p $out
QMCB_BENCH_THERM=3 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
  --log-file $out/r2_launches_full.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --configs lih-opt --no-strong \
  > $out/r2_ncu_bench.log 2>&1
cap() {  # name regex skip script args...
  name=$1; regex=$2; skip=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$regex --launch-skip $skip -c 1 -f \
    -o $out/r2_$name "$@" > $out/r2_${name}.log 2>&1
  ncu -i $out/r2_$name.ncu-rep --page raw --csv > $out/r2_${name}_raw.csv 2>/dev/null
  ncu -i $out/r2_$name.ncu-rep --page source --csv > $out/r2_${name}_src.csv 2>/dev/null
  rm -f $out/r2_$name.ncu-rep
}
cap lih_eloc 'spec_eloc' 2 python tools/profile_eloc.py lih 1000000
cap h2o_eloc 'spect_eloc' 2 python tools/profile_eloc.py h2o 100000
cap c4h6_eloc 'spect_eloc' 2 python tools/profile_eloc.py c4h6 200000
cap c4h6_psi 'spect_psi' 2 python tools/profile_eloc.py c4h6 200000
cap h2o_een_eloc 'spect_eloc' 1 python tests/tools/gpu_config4.py 250000 "cas(4,4)"
cap lih_bwd_spec 'spec_backward' 2 python tools/profile_backward.py lih 1000000
cap lih_bwd_tile 'backward_kernel' 2 python tools/profile_backward.py lih 1000000
ls -la $out/r2_*
