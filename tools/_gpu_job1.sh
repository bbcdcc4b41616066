set -x
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 100 --csv --log-file gpurun_out/r24_launches_c3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r24_ncu_launch.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r24_bench_c3.json 2> gpurun_out/r24_bench_c3.err
