python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python scripts/profile_ops.py > gpurun_out/ops_e.log 2>&1; head -16 gpurun_out/ops_e.log; tail -2 gpurun_out/ops_e.log
