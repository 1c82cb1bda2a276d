mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fem_step -s 10 -c 1 -f -o gpurun_out/r02m_fem_mesh python tools/fem_prof_run.py 148 3 > gpurun_out/r02m_ncu_fem_mesh.log 2>&1
ncu -i gpurun_out/r02m_fem_mesh.ncu-rep --page raw --csv > gpurun_out/r02m_fem_mesh_raw.csv 2>/dev/null
ncu -i gpurun_out/r02m_fem_mesh.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/r02m_fem_mesh_source.csv 2>/dev/null
tail -2 gpurun_out/r02m_ncu_fem_mesh.log; ls -la gpurun_out/r02m_*
