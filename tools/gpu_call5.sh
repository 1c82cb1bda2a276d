mkdir -p gpurun_out
for v in cur; do
  unset TACEX_B200_LIB
  timeout 600 python tools/kbench.py --phases --tag $v > gpurun_out/r02j_kbench_$v.json 2> gpurun_out/r02j_kbench_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02j_kbench_$v.json"))
print(d["tag"], d.get("parity"), d["sparse_fps"], d["dense_fps"], d["box_fps"]); print(d["sparse_phases"]); print(d["dense_phases"]); print(d["sparse_levels"]); print(d["dense_levels"])
PY
  tail -3 gpurun_out/r02j_kbench_$v.err
done
