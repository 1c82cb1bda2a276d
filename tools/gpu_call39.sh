mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_fem_gpu.py -m gpu -x -q -k "two_sided and 3" > gpurun_out/r02o_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|hazard|passed|failed" gpurun_out/r02o_racecheck.log | sort | uniq -c | head -10
