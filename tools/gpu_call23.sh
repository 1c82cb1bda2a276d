mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_shard_gpu.py -m gpu -x -q > gpurun_out/r02ac_shard_pytest.log 2>&1; tail -3 gpurun_out/r02ac_shard_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-fem --no-extras > gpurun_out/r02ac_bench_n2.json 2> gpurun_out/r02ac_bench_n2.err; echo rc=$?
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02ac_bench_n2.json").read().splitlines() if l.startswith("{")][-1])
print(round(d["value"]), round(d["ms_per_step"],2), "verified", d.get("gather_verified"), d["config"]["obs_gather"][:70])
PY
