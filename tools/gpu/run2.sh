set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python tools/ab_kernels.py --out gpurun_out/r2_ab2.jsonl 2>&1 | tail -12
