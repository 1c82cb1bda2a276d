mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_taxim_gpu.py tests/test_lowres.py tests/test_shadow_gpu.py tests/test_overlay.py tests/test_heightmap_gpu.py -m gpu -x -q > gpurun_out/r02ab_pytest.log 2>&1; tail -4 gpurun_out/r02ab_pytest.log
timeout 600 python bench.py --no-fem --no-cpu-baseline > gpurun_out/r02ab_bench.json 2> gpurun_out/r02ab_bench.err; echo rc=$?
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02ab_bench.json").read().splitlines() if l.startswith("{")][-1])
print("value", round(d["value"]), d["ms_per_step"], "dense", round(d["value_dense"]["value"]), "c3box", round(d["value_config3_box"]["value"]), "c2", d.get("value_config2"), "e2e", round(d["e2e"]["value"]))
PY
timeout 200 python tools/phase_times.py 592 0 sparse > gpurun_out/r02ab_phase_sparse.log 2>&1; head -3 gpurun_out/r02ab_phase_sparse.log; tail -3 gpurun_out/r02ab_phase_sparse.log
