mkdir -p gpurun_out
for ch in 576 480 384 288 192; do
  echo "== chunk $ch" >> gpurun_out/r02z_fem_chunk.log
  TX_FEM_CHUNK=$ch timeout 300 python tools/fem_time.py 4096 6 2>&1 | tail -3 >> gpurun_out/r02z_fem_chunk.log
  TX_FEM_CHUNK=$ch timeout 300 ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:fem_step -s 3 -c 1 python tools/fem_prof_run.py 148 2>&1 | grep -E "dram__|gpu__time" >> gpurun_out/r02z_fem_chunk.log
done
cat gpurun_out/r02z_fem_chunk.log
TX_FEM_CHUNK=384 timeout 600 python -m pytest tests/test_fem_gpu.py -m gpu -x -q -k "600 or press_30 or mesh" 2>&1 | tail -3
