python -m pytest tests/test_gpu_peer_gather.py -m gpu -q 2>&1 | tail -3
bash tools/_gpu_multi.sh 2 2>&1 | grep -v "OMP_NUM\|\*\*\*\*\|destroy_process" | cut -c1-900
