set -x
timeout 900 python -m pytest tests/test_k3_integral_gpu.py tests/test_abi.py -m gpu -x -q 2>&1 | tail -15
