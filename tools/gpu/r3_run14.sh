# round 2, session 2, run 14: final validation — GPU suite, smoke, the driver's bench commands (both arms)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference 2>/dev/null | grep '^{' > gpurun_out/r02_v6_bench_reference.json
timeout 600 python bench.py 2>gpurun_out/r02_v6_bench.err | grep '^{' > gpurun_out/r02_v6_bench.json
python - <<'PY'
import json
a=json.load(open('gpurun_out/r02_v6_bench.json')); b=json.load(open('gpurun_out/r02_v6_bench_reference.json'))
print('ours', round(a['value'],1), round(a['ms_per_step'],4), 'e2e', round(a['e2e']['value'],1), 'blocking', round(a['e2e']['blocking_call']['value'],1), 'frac', round(a['roofline']['frac'],4), 'issue', round(a['roofline']['issue']['frac'],4), 'launches', a['gpu_launches'], a['clocks'])
print('ref ', round(b['value'],1), b['cpu_baseline']['cores'], 'same_config', a['config']==b['config'])
print('ratio value', round(a['value']/b['value'],1), 'e2e', round(a['e2e']['value']/b['value'],1))
PY
