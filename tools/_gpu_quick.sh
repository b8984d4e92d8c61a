timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 300 python tools/quick_bench.py 256 256 50
