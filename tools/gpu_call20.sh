mkdir -p gpurun_out
TX_FEM_CHUNK=384 timeout 600 python -m pytest tests/test_fem_gpu.py -m gpu -q -s -k "600 or press_30 or mesh" 2>&1 | grep -E "max \||assert|Error|passed|failed|mismatch" | head -20
