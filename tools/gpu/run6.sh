set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for f in "" "--no-overlap" "--sim-shard 8" "--sim-shard 8 --no-overlap" "--sim-shard 4" "--sim-shard 4 --no-overlap" "--sim-shard 2" "--sim-shard 2 --no-overlap"; do
  timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-e2e $f 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('OVL [$f]', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), d['roofline']['kernel_ms_split'])"
done
timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu > gpurun_out/r2_v3_bench.json 2> gpurun_out/r2_v3_bench.err; tail -c 500 gpurun_out/r2_v3_bench.json
