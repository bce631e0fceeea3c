python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python scripts/profile_ops.py > gpurun_out/ops_x2d.log 2>&1; head -14 gpurun_out/ops_x2d.log; tail -2 gpurun_out/ops_x2d.log
GCCVAE_MARKERS=1 python scripts/graph_timeline.py > gpurun_out/gt_x2d.log 2>&1
grep "main" gpurun_out/gt_x2d.log
