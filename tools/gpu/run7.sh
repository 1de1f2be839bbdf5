set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for f in "" "--overlap" "--sim-shard 8" "--sim-shard 8 --no-overlap" "--sim-shard 4" "--sim-shard 2" "--sim-shard 2 --overlap"; do
  timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-e2e $f 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('OVL [$f]', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), d['roofline']['kernel_ms_split'])"
done
for r in 16 20 24 28; do
  timeout 300 python bench.py --workload picker --steps 5 --warmup 2 --skip-cpu --skip-e2e --refill $r 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('PICK refill $r', round(d['value'],1), round(d['ms_per_step'],4))"
done
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
