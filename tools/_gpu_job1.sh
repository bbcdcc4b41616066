set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r17_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r17_bench_c3.json 2> gpurun_out/r17_bench_c3.err
tail -n 12 gpurun_out/r17_pytest.log
