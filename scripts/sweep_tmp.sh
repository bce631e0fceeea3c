python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python scripts/profile_ops.py 2>&1 | grep "graphs=True"
