mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_pytest.log; tail -3 gpurun_out/r02g_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02g_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02g_smoke.log; tail -2 gpurun_out/r02g_smoke.log
timeout 900 python bench.py > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; echo "bench rc=$?"
TX_TP=1 timeout 300 python tools/fem_time.py 1184 10 4 > gpurun_out/r02g_fem_time_ico_full.log 2>&1; tail -1 gpurun_out/r02g_fem_time_ico_full.log
TX_TP=1 timeout 300 python tools/fem_time.py 4096 12 3 > gpurun_out/r02g_fem_time_cone_full.log 2>&1; tail -3 gpurun_out/r02g_fem_time_cone_full.log
TX_TP=1 timeout 300 python tools/fem_time.py 4096 12 2 2>&1 | tail -1
head -c 300 gpurun_out/r02g_bench.json
