set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r11_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r11_bench_c3.json 2> gpurun_out/r11_bench_c3.err
timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/r11_bench_c5.json 2> gpurun_out/r11_bench_c5.err
timeout 300 python bench.py --workload c1 --steps 20 --warmup 3 > gpurun_out/r11_bench_c1.json 2> gpurun_out/r11_bench_c1.err
timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 > gpurun_out/r11_bench_c2.json 2> gpurun_out/r11_bench_c2.err
timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/r11_bench_c4.json 2> gpurun_out/r11_bench_c4.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_ssao|k_minmax|k_pcss_visibility' -c 6 -o gpurun_out/r11_passes -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu11.log 2>&1
tail -n 5 gpurun_out/r11_pytest.log
