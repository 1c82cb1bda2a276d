mkdir -p gpurun_out
for f in 0 8; do
  timeout 600 python tools/kbench.py --phases --flags $f --tag dyn_flags$f > gpurun_out/r02m_flags$f.json 2> gpurun_out/r02m_flags$f.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02m_flags$f.json"))
print(d["tag"], d.get("parity"), d["sparse_fps"], d["dense_fps"], d["box_fps"]); print(d["sparse_phases"]); print(d["dense_phases"])
PY
  tail -2 gpurun_out/r02m_flags$f.err
done
