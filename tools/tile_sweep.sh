#!/bin/bash
# Launch-shape / option sweep of the warp-tile kernels (spec_tile.cuh) on one workload:
#   bash tools/tile_sweep.sh h2o 100000      |     bash tools/tile_sweep.sh c4h6 20000
key=$1; nw=$2
run() { echo -n "$1 :: "; env $1 QMCB_JIT_CACHE=0 python tools/time_kernels.py $key $nw 2>&1 | tail -1 | sed 's/.*| //'; }
run "QMCB_X=default"
run "QMCB_TILE_MOW=2"
run "QMCB_TILE_MINB=2"
run "QMCB_TILE_MINB=3"
run "QMCB_TILE_MINB=4"
run "QMCB_TILE_MINB=5"
run "QMCB_TILE_THREADS=64 QMCB_TILE_MINB=8"
run "QMCB_TILE_THREADS=64 QMCB_TILE_MINB=6"
run "QMCB_TILE_THREADS=256 QMCB_TILE_MINB=2"
run "QMCB_TILE_THREADS=256 QMCB_TILE_MINB=1"
run "QMCB_SPEC_DEFS=-DSPEC_TILE_PREFETCH=0"
run "QMCB_JIT=0"
