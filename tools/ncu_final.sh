#!/bin/bash
# End-of-round captures of the three E_L kernels the bench line cites (run on a B200 box from the repository
# root): `ncu --set full --clock-control none`, raw + source pages exported as CSV into gpurun_out/.
out=gpurun_out
mkdir -p $out
cap() {  # name regex skip script args...
  name=$1; regex=$2; skip=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$regex --launch-skip $skip -c 1 -f \
    -o /tmp/r2f_$name "$@" > $out/r2f_${name}.log 2>&1
  ncu -i /tmp/r2f_$name.ncu-rep --page raw --csv > $out/r2f_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/r2f_$name.ncu-rep --page source --csv > $out/r2f_${name}_src.csv 2>/dev/null
}
cap lih_eloc 'spec_eloc' 2 python tools/profile_eloc.py lih 1000000
cap h2_eloc 'spec_eloc' 2 python tools/profile_eloc.py h2 1000000
cap h2o_een_eloc 'spect_eloc' 1 python tests/tools/gpu_config4.py 250000 "cas(4,4)"
cap c4h6_eloc 'spect_eloc' 2 python tools/profile_eloc.py c4h6 200000
ls -la $out/r2f_*
