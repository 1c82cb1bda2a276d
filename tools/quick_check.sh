#!/bin/bash
# quick GPU iteration: bit-exact parity tests of the fused kernel + phase clocks + kernel time (extra args: debug flag values)
timeout 600 python -m pytest tests/test_taxim_gpu.py -x -q -k "bitwise or region or camera or batch or fots or full_size" 2>&1 | tail -3
for f in 0 "$@"; do
echo "flags $f"
python tools/phase_times.py 592 $f 2>&1 | awk 'NR<=2 || /epilogue|mask pass/ {print} /hpass/ {h+=$3} /cluster.sync/ {c+=$3} /vpass/ {v+=$3} /reimpose/ {r+=$3} END {print "hpass",h,"csync",c,"vpass",v,"reimpose",r}'
done
