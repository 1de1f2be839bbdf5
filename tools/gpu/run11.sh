cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_group.py -m gpu -x -q -k "two_devices" 2>&1 | tail -60
