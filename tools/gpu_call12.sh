mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_shard_gpu.py -m gpu -x -q > gpurun_out/r02s_shard_pytest.log 2>&1; tail -3 gpurun_out/r02s_shard_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02s_bench_n2.json 2> gpurun_out/r02s_bench_n2.err
echo "rc=$?"; tail -5 gpurun_out/r02s_bench_n2.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02s_bench_n2.json"))
    print({k:d[k] for k in ("value","ms_per_step","n_gpus","gather_verified")}, d["e2e"]["value"], d.get("config4"))
except Exception as e: print("parse failed", e)
PY
