#!/bin/bash
# The per-change GPU check of this repo, run on a B200 box from the repository root:
#   bash tools/gpu_round.sh [tag]
# parity tests, smoke(), kernel timings of the four fixture systems, the bench line and (optional,
# NCU=1) the launch list plus one `ncu --set full` capture of the dominant kernel, exported as CSV.
# Everything lands in gpurun_out/<tag>_*.
tag=${1:-round}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/${tag}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1
( for w in "lih 1000000" "h2 1000000" "lih_sto 1000000" "h2o 100000" "c4h6 20000"; do
    timeout 300 python tools/time_kernels.py $w 2>&1 | tail -1
  done
  timeout 300 python tests/tools/gpu_config4.py 2>&1 | tail -3 ) > $out/${tag}_time.log 2>&1
timeout 300 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2>> $out/${tag}_bench_n1.err
if [ "${NCU:-0}" = "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/${tag}_launches_full.csv python bench.py --steps 5 --warmup 3 --therm 3 --no-cpu-baseline \
    > $out/${tag}_ncu_bench.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:spec_ --launch-skip 6 -c 2 -f \
    -o /tmp/${tag}_spec python tools/profile_eloc.py lih 1000000 > $out/${tag}_ncu.log 2>&1
  ncu -i /tmp/${tag}_spec.ncu-rep --page raw --csv > $out/${tag}_spec_raw.csv 2>/dev/null
fi
tail -3 $out/${tag}_gpu_tests.log; tail -2 $out/${tag}_smoke.log; cat $out/${tag}_time.log; cat $out/${tag}_bench_n1.json
