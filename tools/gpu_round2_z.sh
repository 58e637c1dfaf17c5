set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --tb=short -p no:cacheprovider -k "free_dt" 2>&1 | tail -5
