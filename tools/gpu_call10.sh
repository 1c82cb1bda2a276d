mkdir -p gpurun_out
export TACEX_B200_LIB=$PWD/tacex_b200/lib/libtacex_b200_vstat.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:taxim_fused -s 2 -c 1 -f -o gpurun_out/r02q_dyn python tools/prof_run.py 592 sparse > gpurun_out/r02q_ncu.log 2>&1
tail -2 gpurun_out/r02q_ncu.log
