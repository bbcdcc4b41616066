set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r20_pytest.log
tail -n 12 gpurun_out/r20_pytest.log
