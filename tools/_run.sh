python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1c_n1.json 2> gpurun_out/bench_r1c_n1.err; cut -c1-300 gpurun_out/bench_r1c_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 5 --warmup 3 --therm 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_r1c.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:spec_eloc|spec_psi|spec_mh" -s 5 -c 3 -o gpurun_out/spec_r1d -f python tools/profile_eloc.py lih 1000000 > /dev/null 2>&1
ls -la gpurun_out/spec_r1d.ncu-rep gpurun_out/launches_r1c.csv
for k in h2 h2o c4h6; do python tools/time_kernels.py $k $( [ $k = h2 ] && echo 1000000 || ([ $k = h2o ] && echo 100000 || echo 20000) ) | tail -1; done
python tools/profile_backward.py lih 1000000 2>&1 | tail -3
