set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r06_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r06_bench_c3.json 2> gpurun_out/r06_bench_c3.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_chain_fused' -c 1 -o gpurun_out/r06_chain -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_chain6.log 2>&1
timeout 300 python bench.py --workload c1 --steps 20 --warmup 3 > gpurun_out/r06_bench_c1.json 2> gpurun_out/r06_bench_c1.err
tail -5 gpurun_out/r06_pytest.log; tail -3 gpurun_out/r06_bench_c3.err
