mkdir -p gpurun_out
export TACEX_B200_LIB=$PWD/tacex_b200/lib/libtacex_b200_vstat.so
for f in 64 72; do
  timeout 600 python tools/kbench.py --phases --no-parity --flags $f --tag vstat_flags$f > gpurun_out/r02p_flags$f.json 2> gpurun_out/r02p_flags$f.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02p_flags$f.json"))
print(d["tag"], d["sparse_fps"], d["dense_fps"], d["box_fps"]); print(d["sparse_phases"]); print(d["sparse_levels"])
PY
done
