python -m pytest tests/test_gpu_x2.py -x -q 2>&1 | tail -1
python scripts/timeline_probe.py 2>&1 | grep "^==\|back-to-back" | grep -A1 "C3CONV" | grep -v "^--" | cut -c1-60
python scripts/profile_ops.py 2>&1 | grep "graphs=True"
