set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_rust_shim.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
for f in "" "--ctas-per-sm 6" "--ctas-per-sm 5" "--ctas-per-sm 4" "--ctas-per-sm 3"; do
  timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-e2e --sim-shard 8 $f 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('SIM8 [$f]', round(d['ms_per_step'],4), d['roofline']['kernel_ms'], d['roofline']['kernel_ms_split'], d['gpu_launches'], d['host_issue_ms_per_step'])"
done
timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-e2e --sim-shard 8 --no-flush 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('SIM8 noflush', round(d['ms_per_step'],4), d['roofline']['kernel_ms'], d['roofline']['kernel_ms_split'])"
timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu > gpurun_out/r2_v2_bench.json 2> gpurun_out/r2_v2_bench.err; tail -c 700 gpurun_out/r2_v2_bench.json
