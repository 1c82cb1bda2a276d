mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fem_gpu.py -m gpu -x -q -s -k "4096" > gpurun_out/r02k_fem4096.log 2>&1; tail -6 gpurun_out/r02k_fem4096.log
