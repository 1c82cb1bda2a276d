mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_fem_gpu.py -m gpu -x -q -s > gpurun_out/r02ai_fem_pytest.log 2>&1; tail -9 gpurun_out/r02ai_fem_pytest.log
TX_TP=1 timeout 300 python tools/fem_time.py 1184 10 4 > gpurun_out/r02ai_fem_time_ico_full.log 2>&1; tail -3 gpurun_out/r02ai_fem_time_ico_full.log
TX_TP=1 timeout 300 python tools/fem_time.py 4096 12 3 2>&1 | tail -1
timeout 300 python tools/fem_time.py 4096 6 2>&1 | tail -1
