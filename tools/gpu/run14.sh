cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python tools/gpu/stress_group.py 400 2>&1 | grep -v "NCCL\|Exception\|Traceback\|File\|TypeError"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
