mkdir -p gpurun_out
for v in cur vstat; do
if [ $v = cur ]; then unset TACEX_B200_LIB; else export TACEX_B200_LIB=$PWD/tacex_b200/lib/libtacex_b200_$v.so; fi
for f in 0 8; do
  timeout 600 python tools/kbench.py --phases --flags $f --tag ${v}_flags$f > gpurun_out/r02o_${v}_flags$f.json 2> gpurun_out/r02o_${v}_flags$f.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02o_${v}_flags$f.json"))
print(d["tag"], d.get("parity",{}).get("rgb"), d["sparse_fps"], d["dense_fps"], d["box_fps"]); print(d["sparse_phases"]); print(d["dense_phases"])
PY
  tail -2 gpurun_out/r02o_${v}_flags$f.err
done
done
