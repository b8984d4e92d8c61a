timeout 900 python -m pytest tests/test_gpu_integrate.py -m gpu -q 2>&1 | tail -8
timeout 600 python tools/diag_integrate.py f3d_s256_view5 unit_p20000_sh3_bg 2>&1 | grep -E "==|time"
