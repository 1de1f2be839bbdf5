set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for f in "" "--sim-shard 8"; do
  timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-e2e $f 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('CLIP [$f]', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), d['roofline']['kernel_ms_split'], d['roofline']['counts']['steps'])"
done
timeout 300 python bench.py --workload picker --steps 5 --warmup 2 --skip-cpu --skip-e2e --refill 20 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('PICK', round(d['value'],1), round(d['ms_per_step'],4), d['roofline']['counts'])"
timeout 300 python bench.py --format csvo --steps 20 --warmup 5 --skip-cpu --skip-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('CSVO', round(d['value'],1), round(d['ms_per_step'],4), d['roofline']['kernel_ms_split'])"
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_v4_bench.json 2> gpurun_out/r2_v4_bench.err; tail -c 600 gpurun_out/r2_v4_bench.json
