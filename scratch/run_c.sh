timeout 600 python -m pytest tests/test_gpu_abi_client.py tests/test_gpu_parity.py -m gpu -q -x -k "c_client or find_c1 or small" 2>&1 | tail -8
