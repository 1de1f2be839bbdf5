set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/ab_kernels.py --out gpurun_out/r2_ab3.jsonl --variants stateasm,popidx 2>&1 | tail -5
for f in "--ctas-per-sm 6" "--ctas-per-sm 7" "--refill 2" "--refill 4" "--refill-shadow 2"; do
  timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-e2e $f 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$f', round(d['value'],1), round(d['ms_per_step'],4), d['roofline']['kernel_ms_split'])"
done
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
