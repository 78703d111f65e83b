#!/bin/bash
out=gpurun_out/r02_agknobs.txt
: > $out
run() { echo "== $*" >> $out; for r in 1 2; do env "$@" timeout 120 python scripts/kernel_bench.py --n 24 26 --path 4 --steps 8 2>&1 | cut -c1-120 >> $out; done; }
run PD_STREAM_HINTS=0
run PD_STREAM_HINTS=3
run PD_STREAM_HINTS=3 PD_STREAM_CHUNK=7 PD_STREAM_LAG=2
run PD_STREAM_HINTS=3 PD_STREAM_CHUNK=7 PD_STREAM_LAG=3 PD_STREAM_MIX=1
run PD_STREAM_HINTS=3 PD_STREAM_G1BITS=6 PD_STREAM_CHUNK=6 PD_STREAM_LAG=4
run PD_STREAM_HINTS=3 PD_STREAM_G1BITS=6 PD_STREAM_CHUNK=6 PD_STREAM_LAG=6 PD_STREAM_MIX=1
run PD_STREAM_HINTS=3 PD_STREAM_G1BITS=6 PD_STREAM_CHUNK=6 PD_STREAM_LAG=8 PD_STREAM_MIX=1
run PD_STREAM_HINTS=7 PD_STREAM_G1BITS=6 PD_STREAM_CHUNK=6 PD_STREAM_LAG=6 PD_STREAM_MIX=1
echo "== FUSE=2 test variants" >> $out
for cfg in "PD_STREAM_GPIPE=0" "PD_STREAM_GPIPE=0 PD_STREAM_TMA_G=1" "PD_STREAM_GPIPE=1 PD_STREAM_TMA_G=1"; do
 echo "-- $cfg" >> $out
 env $cfg PD_STREAM_FUSE=2 timeout 300 python -m pytest tests/test_gpu_scale.py -m gpu -q -k "test_tiled_equals_gather and (26-4 or 24-4 or 25-4)" 2>&1 | tail -4 >> $out
done
cat $out
