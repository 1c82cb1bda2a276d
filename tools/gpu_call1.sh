mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r02a_gpu.log 2>&1
nproc >> gpurun_out/r02a_gpu.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02a_smoke.log
timeout 900 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?" >> gpurun_out/r02a_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02a_bench_ref.json 2> gpurun_out/r02a_bench_ref.err
timeout 200 python tools/phase_times.py 592 0 sparse > gpurun_out/r02a_phase_sparse.log 2>&1
timeout 200 python tools/phase_times.py 592 0 dense > gpurun_out/r02a_phase_dense.log 2>&1
tail -3 gpurun_out/r02a_pytest.log; tail -2 gpurun_out/r02a_smoke.log; head -c 600 gpurun_out/r02a_bench.json; tail -3 gpurun_out/r02a_bench.err
