set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r13_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r13_bench_c3.json 2> gpurun_out/r13_bench_c3.err
timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r13_bench_c5.json 2> gpurun_out/r13_bench_c5.err
tail -n 5 gpurun_out/r13_pytest.log
