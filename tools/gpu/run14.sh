cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python tools/gpu/stress_group.py 300 2>&1 | grep -v NCCL
