mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fem_gpu.py -m gpu -x -q -s > gpurun_out/r02w_fem_pytest.log 2>&1; tail -12 gpurun_out/r02w_fem_pytest.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/hess12_dmma tools/ubench/hess12_dmma.cu && timeout 120 /tmp/hess12_dmma > gpurun_out/r02w_hess12_dmma.json; cat gpurun_out/r02w_hess12_dmma.json
timeout 300 python tools/fem_time.py 4096 6 > gpurun_out/r02w_fem_time_box.log 2>&1; tail -3 gpurun_out/r02w_fem_time_box.log
timeout 300 python tools/fem_time.py 4096 12 2 > gpurun_out/r02w_fem_time_wedge.log 2>&1; tail -4 gpurun_out/r02w_fem_time_wedge.log
timeout 300 python tools/fem_time.py 4096 12 3 > gpurun_out/r02w_fem_time_cone.log 2>&1; tail -4 gpurun_out/r02w_fem_time_cone.log
