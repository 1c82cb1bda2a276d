mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fem_gpu.py tests/test_taxim_gpu.py -m gpu -x -q > gpurun_out/r02ad_pytest.log 2>&1; tail -4 gpurun_out/r02ad_pytest.log
timeout 300 python tools/fem_time.py 4096 6 2>&1 | tail -2
timeout 600 python bench.py --no-fem --no-cpu-baseline > gpurun_out/r02ad_bench.json 2> gpurun_out/r02ad_bench.err; echo rc=$?
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02ad_bench.json").read().splitlines() if l.startswith("{")][-1])
print("value", round(d["value"]), d["ms_per_step"], "dense", round(d["value_dense"]["value"]), "c3box", round(d["value_config3_box"]["value"]), "c2", round(d["value_config2"]["value"]), "e2e", round(d["e2e"]["value"]))
PY
