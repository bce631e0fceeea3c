python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/profile_ops.py > gpurun_out/ops_s2d.log 2>&1; head -20 gpurun_out/ops_s2d.log; tail -2 gpurun_out/ops_s2d.log
GCCVAE_S2D=0 python scripts/profile_ops.py 2>&1 | tail -1
