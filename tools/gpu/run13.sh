cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_group.py -m gpu -x -q -k "two_devices" 2>&1 | grep -E "passed|failed|AssertionError:" ; done
timeout 600 python tools/gpu/dbg_group.py 2>&1 | grep -c "bad px 0"
