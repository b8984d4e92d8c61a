timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py -m gpu -q -x -k "forward or batch or workspace or render_views or big_tile" 2>&1 | tail -60
