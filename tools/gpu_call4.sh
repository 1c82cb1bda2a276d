mkdir -p gpurun_out
for k in sparse dense; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:taxim_fused -s 2 -c 1 -f -o gpurun_out/r02d_$k python tools/prof_run.py 592 $k > gpurun_out/r02d_ncu_$k.log 2>&1
tail -2 gpurun_out/r02d_ncu_$k.log
done
ls -la gpurun_out/*.ncu-rep
