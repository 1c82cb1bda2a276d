mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_fem_gpu.py -m gpu -x -q -k "two_sided" > gpurun_out/r02n_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|passed|failed|error" gpurun_out/r02n_memcheck.log | head -10
