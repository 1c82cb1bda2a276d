mkdir -p gpurun_out
for v in base cur; do
  if [ $v = cur ]; then unset TACEX_B200_LIB; else export TACEX_B200_LIB=$PWD/tacex_b200/lib/libtacex_b200_$v.so; fi
  timeout 600 python tools/kbench.py --phases --tag $v > gpurun_out/r02b_kbench_$v.json 2> gpurun_out/r02b_kbench_$v.err
  tail -c 1500 gpurun_out/r02b_kbench_$v.json; tail -3 gpurun_out/r02b_kbench_$v.err
done
