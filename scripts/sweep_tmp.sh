python -m pytest tests/test_gpu_x2.py tests/test_gpu_parity_bf16.py -x -q 2>&1 | tail -2
python scripts/profile_ops.py 2>&1 | grep "conv1 wgrad\|conv5t wgrad\|graphs=True"
