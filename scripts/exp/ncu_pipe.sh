#!/bin/bash
PD_STREAM_FUSE=2 PD_STREAM_HINTS=3 PD_STREAM_G1BITS=6 PD_STREAM_CHUNK=6 PD_STREAM_LAG=4 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_stream_(pipe|ag)" -s 3 -c 1 \
  -o gpurun_out/r02_pipe_v3 -f python scripts/kernel_bench.py --n 26 --path 4 --steps 1 > gpurun_out/ncu_pipe_v3.log 2>&1
ls -la gpurun_out/*.ncu-rep
