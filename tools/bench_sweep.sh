#!/bin/bash
# Launch-shape sweep of the specialised E_L kernel through bench.py itself (cold-L2 ensembles):
#   bash tools/bench_sweep.sh            (on a B200 box; prints value, ms/step, kernel ms, roofline.frac)
run() { echo "== $*"; env "$@" timeout 120 python bench.py --no-cpu-baseline --therm 20 $EXTRA 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.4e ms/step %.4f kernel_ms %.4f frac %.4f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))"; }
if [ "$1" = "second" ]; then
  run A=0
  run QMCB_SPEC_THREADS=160 QMCB_SPEC_MINB=3 QMCB_SPEC_DEFS=-DSPEC_MINB_ELOC=3
  run QMCB_SPEC_THREADS=96 QMCB_SPEC_MINB=5 QMCB_SPEC_DEFS=-DSPEC_MINB_ELOC=5
  run QMCB_SPEC_THREADS=192 QMCB_SPEC_MINB=2 QMCB_SPEC_DEFS=-DSPEC_MINB_ELOC=2
  run QMCB_SPEC_THREADS=64 QMCB_SPEC_MINB=6 QMCB_SPEC_DEFS=-DSPEC_MINB_ELOC=5
  EXTRA="--walkers 1022976" run A=0
  exit 0
fi
run A=0
run QMCB_SPEC_THREADS=64 QMCB_SPEC_MINB=6 QMCB_SPEC_DEFS=-DSPEC_MINB_ELOC=6
run QMCB_SPEC_THREADS=96 QMCB_SPEC_MINB=4 QMCB_SPEC_DEFS=-DSPEC_MINB_ELOC=4
run QMCB_SPEC_THREADS=64 QMCB_SPEC_MINB=6 QMCB_SPEC_DEFS=-DSPEC_MINB_ELOC=7
run QMCB_SPEC_THREADS=32 QMCB_SPEC_MINB=12 QMCB_SPEC_DEFS=-DSPEC_MINB_ELOC=12
run QMCB_STATS_2STAGE=1
run QMCB_SPEC_THREADS=256 QMCB_SPEC_MINB=1 QMCB_SPEC_DEFS=-DSPEC_MINB_ELOC=1
run A=0
