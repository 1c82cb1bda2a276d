mkdir -p gpurun_out
N=${NG:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 tools/pcie_probe.py 2> gpurun_out/r02t_pcie.err | grep '^{' > gpurun_out/r02t_pcie_n$N.json; cat gpurun_out/r02t_pcie_n$N.json
timeout 600 $TR --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02t_bench_n${N}_fp32.json 2> gpurun_out/r02t_bench_n${N}_fp32.err; echo rc=$?
timeout 400 $TR --master-port 29523 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-fem --no-extras --obs-gather u8 > gpurun_out/r02t_bench_n${N}_u8.json 2> gpurun_out/r02t_bench_n${N}_u8.err; echo rc=$?
timeout 400 $TR --master-port 29524 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-fem --no-extras --obs-gather none > gpurun_out/r02t_bench_n${N}_none.json 2> gpurun_out/r02t_bench_n${N}_none.err; echo rc=$?
python - <<PY
import json
for m in ("fp32","u8","none"):
    try:
        d=json.loads([l for l in open("gpurun_out/r02t_bench_n${N}_%s.json"%m).read().splitlines() if l.startswith("{")][-1])
        print(m, round(d["value"]), round(d["ms_per_step"],2), "verified", d.get("gather_verified"), "e2e", round(d["e2e"]["value"]), (d.get("config4") or {}).get("frames_per_s"), (d.get("config4") or {}).get("gather_verified"))
    except Exception as e: print(m, "failed", e)
PY
tail -3 gpurun_out/r02t_bench_n${N}_fp32.err
