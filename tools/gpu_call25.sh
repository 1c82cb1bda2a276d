mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log; tail -3 gpurun_out/r02f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02f_smoke.log; tail -2 gpurun_out/r02f_smoke.log
timeout 900 python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02f_bench_ref.json 2> gpurun_out/r02f_bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02f_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:taxim_fused -s 2 -c 1 -f -o gpurun_out/r02f_taxim python tools/prof_run.py 4096 sparse > gpurun_out/r02f_ncu_taxim.log 2>&1
ncu -i gpurun_out/r02f_taxim.ncu-rep --page raw --csv > gpurun_out/r02f_taxim_raw.csv 2>/dev/null
ncu -i gpurun_out/r02f_taxim.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/r02f_taxim_source.csv 2>/dev/null
ls -la gpurun_out/r02f_taxim.ncu-rep
[ $(stat -c %s gpurun_out/r02f_taxim.ncu-rep) -gt 30000000 ] && rm gpurun_out/r02f_taxim.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fem_step -s 3 -c 1 -f -o gpurun_out/r02f_fem python tools/fem_prof_run.py 148 > gpurun_out/r02f_ncu_fem.log 2>&1
ncu -i gpurun_out/r02f_fem.ncu-rep --page raw --csv > gpurun_out/r02f_fem_raw.csv 2>/dev/null
timeout 200 python tools/phase_times.py 592 0 sparse > gpurun_out/r02f_phase_sparse.log 2>&1; tail -3 gpurun_out/r02f_phase_sparse.log
head -c 400 gpurun_out/r02f_bench.json; tail -3 gpurun_out/r02f_bench.err
