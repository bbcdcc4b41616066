set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r21_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r21_bench_c3.json 2> gpurun_out/r21_bench_c3.err
FGL_NO_CHAIN_OVERLAP=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r21_bench_c3_noov.json 2> gpurun_out/r21_bench_c3_noov.err
timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r21_bench_c5.json 2> gpurun_out/r21_bench_c5.err
tail -n 12 gpurun_out/r21_pytest.log
