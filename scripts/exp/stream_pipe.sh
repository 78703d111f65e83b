#!/bin/bash
out=gpurun_out/r02_pipe4.txt
: > $out
for cfg in "PD_STREAM_FUSE=2" "PD_STREAM_FUSE=2 PD_STREAM_GPIPE=0" "PD_STREAM_FUSE=2 PD_STREAM_TMA_G=1"; do
  echo "== stress $cfg" >> $out; env $cfg REPS=400 timeout 200 python scripts/exp/pipe_diag4.py 2>&1 | tail -1 >> $out
done
echo "== tests FUSE=2" >> $out
PD_STREAM_FUSE=2 PD_STREAM_HINTS=3 timeout 600 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "tiled or stream or oracle_sparse or families" 2>&1 | tail -3 >> $out
run() { echo "== $*" >> $out; env "$@" timeout 120 python scripts/kernel_bench.py --n 24 26 --path 4 --steps 4 2>&1 | cut -c1-250 >> $out; }
run PD_STREAM_FUSE=2 PD_STREAM_HINTS=3
run PD_STREAM_FUSE=2 PD_STREAM_HINTS=3 PD_STREAM_GPIPE=0
run PD_STREAM_FUSE=2 PD_STREAM_HINTS=3 PD_STREAM_G1BITS=6 PD_STREAM_CHUNK=6 PD_STREAM_LAG=4
run PD_STREAM_FUSE=2 PD_STREAM_HINTS=3 PD_STREAM_G1BITS=6 PD_STREAM_CHUNK=6 PD_STREAM_LAG=4 PD_STREAM_GPIPE=0
run PD_STREAM_FUSE=2 PD_STREAM_HINTS=3 PD_STREAM_G1BITS=6 PD_STREAM_CHUNK=6 PD_STREAM_LAG=4 PD_STREAM_TMA_G=1
run PD_STREAM_FUSE=1 PD_STREAM_HINTS=3
cat $out
