set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "wide or config4 or large_grid" 2>&1 | tail -5
timeout 300 python tools/dev_timing_wide.py 2>&1 | tail -12
WIDE_QUICK=1 timeout 600 python tools/time_wide.py 2>&1 | tail -40
