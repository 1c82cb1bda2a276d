mkdir -p gpurun_out
for f in 0 1 2 16 32 48 51; do
  timeout 300 python tools/kbench.py --phases --no-parity --flags $f --tag flags$f > gpurun_out/r02k_flags$f.json 2> gpurun_out/r02k_flags$f.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02k_flags$f.json"))
print(d["tag"], d["sparse_fps"], d["dense_fps"], "colour", d["sparse_phases"]["colour"], d["dense_phases"]["colour"], "total", d["sparse_phases"]["total"], d["dense_phases"]["total"])
PY
done
