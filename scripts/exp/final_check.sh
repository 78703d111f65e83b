#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02_final_tests.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err; echo "bench rc=$?" >> gpurun_out/r02_final_tests.txt
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:"k_stream" --clock-control none -c 40 --csv \
  --log-file gpurun_out/r02_traffic_ncu.csv python scripts/kernel_bench.py --n 26 --path 4 --steps 1 > /dev/null 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
cat gpurun_out/r02_final_tests.txt
