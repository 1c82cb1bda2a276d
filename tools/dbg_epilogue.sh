for f in 0 8; do echo "== flags $f"; TX_DEBUG_FLAGS=$f timeout 100 python tools/phase_times.py 592 2>&1 | grep -E "kernel time|epilogue|total cycles"; done
