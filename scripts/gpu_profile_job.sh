set -x
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench exit $?"
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --skip-n23 --roofline-steps 2 > gpurun_out/plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --skip-n23 --roofline-steps 2 > gpurun_out/ncu_bench.log 2>&1; echo "ncu1 exit $?"
timeout 300 python scripts/kernel_bench.py --n 26 --path 4 --steps 2 > gpurun_out/plain_kb.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_n26_stream.csv python scripts/kernel_bench.py --n 26 --path 4 --steps 2 > gpurun_out/ncu_kb.log 2>&1; echo "ncu2 exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream -s 12 -c 3 -o gpurun_out/prof_stream_n26 python scripts/kernel_bench.py --n 26 --path 4 --steps 2 > gpurun_out/ncu_kb2.log 2>&1; echo "ncu3 exit $?"
tail -c 600 gpurun_out/bench_r01.json
timeout 200 python scripts/small_bench.py 12 0 > gpurun_out/plain_small.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_small -c 2 -o gpurun_out/prof_small_final python scripts/small_bench.py 12 0 > gpurun_out/ncu_small_final.log 2>&1; echo "ncu4 exit $?"
