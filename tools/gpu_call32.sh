mkdir -p gpurun_out
N=${NG:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 700 $TR --master-port 29571 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02h_bench_n${N}.json 2> gpurun_out/r02h_bench_n${N}.err; echo rc=$?
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02h_bench_n${N}.json").read().splitlines() if l.startswith("{")][-1])
    c4=d.get("config4") or {}
    print(round(d["value"]), round(d["ms_per_step"],2), "verified", d.get("gather_verified"), "e2e", round(d["e2e"]["value"]), "c4", c4.get("frames_per_s"), c4.get("gather_verified"), "c2", (d.get("value_config2") or {}).get("value"), "dense", (d.get("value_dense") or {}).get("value"))
except Exception as e: print("failed", e)
PY
tail -3 gpurun_out/r02h_bench_n${N}.err
timeout 400 $TR --master-port 29572 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r02h_bench_ref_n${N}.json 2> gpurun_out/r02h_bench_ref_n${N}.err; echo ref rc=$?; head -c 300 gpurun_out/r02h_bench_ref_n${N}.json
