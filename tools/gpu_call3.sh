mkdir -p gpurun_out
for v in cur rec80; do
  if [ $v = cur ]; then unset TACEX_B200_LIB; else export TACEX_B200_LIB=$PWD/tacex_b200/lib/libtacex_b200_$v.so; fi
  timeout 600 python tools/kbench.py --phases --tag $v > gpurun_out/r02c_kbench_$v.json 2> gpurun_out/r02c_kbench_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02c_kbench_$v.json"))
print(d["tag"], d.get("parity"), d["sparse_fps"], d["dense_fps"], d["box_fps"]); print(d["sparse_phases"]); print(d["dense_phases"])
PY
  tail -3 gpurun_out/r02c_kbench_$v.err
done
