cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for f in "" "--morton" "--sim-shard 8" "--sim-shard 8 --no-overlap" "--sim-shard 4" "--sim-shard 4 --no-overlap"; do
  timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-e2e $f 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('R15 [$f]', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), d['roofline']['kernel_ms_split'], d['roofline'].get('l2',{}).get('peak'), d['roofline'].get('l2',{}).get('hbm_read_peak_measured_here'))"
done
