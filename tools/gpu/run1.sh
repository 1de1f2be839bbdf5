set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python tools/ab_kernels.py --out gpurun_out/r2_ab1.jsonl --variants r01,plain 2>&1 | tail -8
timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-e2e --no-l2-window > gpurun_out/r2_no_l2_window.json 2> gpurun_out/r2_no_l2_window.err
timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-e2e --tma > gpurun_out/r2_tma.json 2> gpurun_out/r2_tma.err
timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu > gpurun_out/r2_v1_bench.json 2> gpurun_out/r2_v1_bench.err
tail -c 600 gpurun_out/r2_v1_bench.json
