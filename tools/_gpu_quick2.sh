for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "test_backward and single_gaussian" 2>&1 | grep -E "^E  |passed|failed" | head -5; done
