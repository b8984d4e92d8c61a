timeout 300 python tools/diag_bwd.py f3d_s256_r256_view2 2>&1 | grep -v worst | head -30
