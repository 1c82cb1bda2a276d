mkdir -p gpurun_out
for fl in 0 512 1024 2048; do
  timeout 300 python tools/kbench.py --flags $fl --tag bias$fl --n 4096 > gpurun_out/r02ag_bias$fl.json 2> gpurun_out/r02ag_bias$fl.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02ag_bias$fl.json").read().splitlines() if l.startswith("{")][-1])
print($fl, d.get("parity",{}).get("rgb"), d.get("sparse_fps"), d.get("dense_fps"), d.get("box_fps"))
PY
done
