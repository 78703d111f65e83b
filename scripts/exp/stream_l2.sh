#!/bin/bash
# A/B of the L2-residency knobs of the stream family's dataflow launch (N=26)
out=gpurun_out/r02_l2knobs.txt
: > $out
run() { echo "== $*" >> $out; env "$@" timeout 120 python scripts/kernel_bench.py --n 26 --path 4 --steps 4 2>&1 | cut -c1-140 >> $out; }
run PD_STREAM_HINTS=0
run PD_STREAM_HINTS=1
run PD_STREAM_HINTS=3
run PD_STREAM_HINTS=7
run PD_STREAM_HINTS=0 PD_STREAM_G1BITS=6 PD_STREAM_CHUNK=6 PD_STREAM_LAG=1
run PD_STREAM_HINTS=0 PD_STREAM_G1BITS=6 PD_STREAM_CHUNK=6 PD_STREAM_LAG=3
run PD_STREAM_HINTS=3 PD_STREAM_G1BITS=6 PD_STREAM_CHUNK=6 PD_STREAM_LAG=1
run PD_STREAM_HINTS=3 PD_STREAM_G1BITS=6 PD_STREAM_CHUNK=6 PD_STREAM_LAG=2
run PD_STREAM_HINTS=3 PD_STREAM_G1BITS=6 PD_STREAM_CHUNK=6 PD_STREAM_LAG=4
run PD_STREAM_HINTS=7 PD_STREAM_G1BITS=6 PD_STREAM_CHUNK=6 PD_STREAM_LAG=2
run PD_STREAM_HINTS=3 PD_STREAM_CHUNK=7 PD_STREAM_LAG=2
run PD_STREAM_HINTS=3 PD_STREAM_G1BITS=6 PD_STREAM_CHUNK=8 PD_STREAM_LAG=1
for cfg in "PD_STREAM_HINTS=0" "PD_STREAM_HINTS=3 PD_STREAM_G1BITS=6 PD_STREAM_CHUNK=6 PD_STREAM_LAG=2"; do
  tag=$(echo $cfg | tr ' =' '__')
  env $cfg timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
    -k regex:k_stream --clock-control none -c 14 --csv --log-file gpurun_out/r02_l2knobs_ncu_$tag.csv \
    python scripts/kernel_bench.py --n 26 --path 4 --steps 1 > /dev/null 2>&1
done
cat $out
