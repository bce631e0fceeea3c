python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scripts/profile_ops.py > gpurun_out/ops_x2b.log 2>&1
head -12 gpurun_out/ops_x2b.log; tail -2 gpurun_out/ops_x2b.log
GCCVAE_C3_PER_SM=3 python scripts/profile_ops.py 2>&1 | grep "enc.conv1 fwd\|conv5t dgrad\|graphs="
GCCVAE_MARKERS=1 python scripts/graph_timeline.py > gpurun_out/gt_x2b.log 2>&1
grep "main" gpurun_out/gt_x2b.log | grep "conv1 \|conv5t\|prep\|adam"
