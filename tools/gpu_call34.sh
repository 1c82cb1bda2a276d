mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_fem_gpu.py -m gpu -x -q -s > gpurun_out/r02j_fem_pytest.log 2>&1; tail -9 gpurun_out/r02j_fem_pytest.log
TX_TP=1 timeout 300 python tools/fem_time.py 4096 12 3 2>&1 | tail -1
timeout 300 python tools/fem_time.py 4096 6 2>&1 | tail -1
