# round 2, session 2, run 23: bench lines of the shipped build for profiles/ (frame, picker)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python bench.py 2>/dev/null | grep '^{' > gpurun_out/r02_v6_bench.json
timeout 400 python bench.py --workload picker --steps 8 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/r02_v6_bench_picker.json
python - <<'PY'
import json
for f in ('r02_v6_bench','r02_v6_bench_picker'):
    d=json.load(open('gpurun_out/%s.json'%f)); e=d['e2e']
    print(f, round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(e['value'],1), round(e['ms_per_step'],3), 'cpu', d.get('cpu_baseline',{}).get('value'))
PY
