timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "receding" 2>&1 | tail -5
