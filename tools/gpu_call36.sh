mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02l_bench.json").read().splitlines() if l.startswith("{")][-1])
print(round(d["value"]), round(d["e2e"]["value"]), d.get("value_lowres_32x32"), d["fem_gel_substep"]["mesh_indenters"]["cone"]["gel_steps_per_s"])
PY
tail -2 gpurun_out/r02l_bench.err
