python tests/golden/make_golden.py gpurun_out/golden 2>&1 | tail -5
