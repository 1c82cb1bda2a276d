mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29581 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r02i_bench_ref_n2.json 2> gpurun_out/r02i_bench_ref_n2.err; echo ref rc=$?
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02i_bench_ref_n2.json').read().splitlines() if l.startswith('{')][-1]); print(d['value'], d['cpu_baseline']['cores'], d['n_gpus'])"
