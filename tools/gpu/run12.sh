cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/gpu/dbg_group.py 2>&1 | grep -v NCCL | head -24
