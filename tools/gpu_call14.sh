mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_shard_gpu.py -m gpu -x -q > gpurun_out/r02u_shard_pytest.log 2>&1; tail -4 gpurun_out/r02u_shard_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for m in fp32-fused fp32-rect; do
timeout 400 $TR --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-fem --no-extras --obs-gather $m > gpurun_out/r02u_bench_n2_$m.json 2> gpurun_out/r02u_bench_n2_$m.err; echo rc=$?
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02u_bench_n2_$m.json").read().splitlines() if l.startswith("{")][-1])
    print("$m", round(d["value"]), round(d["ms_per_step"],2), "verified", d.get("gather_verified"), d["config"]["obs_gather"][:90])
except Exception as e: print("$m failed", e)
PY
tail -3 gpurun_out/r02u_bench_n2_$m.err
done
