mkdir -p gpurun_out
for p in 0 1; do
  TX_FEM_L2_PERSIST=$p timeout 300 python tools/fem_time.py 4096 6 > gpurun_out/r02r_fem_persist$p.log 2>&1
  tail -3 gpurun_out/r02r_fem_persist$p.log
done
TX_FEM_L2_PERSIST=1 timeout 600 ncu --set full --clock-control none -k regex:fem_step -s 3 -c 1 -f -o gpurun_out/r02r_fem_persist1 python tools/fem_prof_run.py 148 > gpurun_out/r02r_ncu_fem.log 2>&1
tail -2 gpurun_out/r02r_ncu_fem.log
