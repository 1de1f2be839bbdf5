set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 900 python tools/ab_kernels.py --out gpurun_out/r2_ab9.jsonl --variants noprimclip,hdr128 2>&1 | tail -5
for f in "--morton" "--no-clip" "--no-l2-window"; do
  timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-e2e $f 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('AB [$f]', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), d['roofline']['kernel_ms_split'], d['roofline']['counts']['steps'])"
done
timeout 300 python bench.py --workload picker --steps 5 --warmup 2 --skip-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('PICK', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))"
