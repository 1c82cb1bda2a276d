mkdir -p gpurun_out
N=${NG:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02v_bench_n${N}_fused.json 2> gpurun_out/r02v_bench_n${N}_fused.err; echo rc=$?
timeout 300 $TR --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-fem --no-extras --obs-gather fp32-rect > gpurun_out/r02v_bench_n${N}_rect.json 2> gpurun_out/r02v_bench_n${N}_rect.err; echo rc=$?
python - <<PY
import json
for m in ("fused","rect"):
    try:
        d=json.loads([l for l in open("gpurun_out/r02v_bench_n${N}_%s.json"%m).read().splitlines() if l.startswith("{")][-1])
        print(m, round(d["value"]), round(d["ms_per_step"],2), "verified", d.get("gather_verified"), "e2e", round(d["e2e"]["value"]), (d.get("config4") or {}).get("frames_per_s"), (d.get("config4") or {}).get("gather_verified"), d["config"]["obs_gather"][:80])
    except Exception as e: print(m, "failed", e)
PY
tail -3 gpurun_out/r02v_bench_n${N}_fused.err
