# round 2, session 2, run 11: bench lines of the final build (not under a profiler)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --steps 30 --warmup 5 2>gpurun_out/r02_v6_bench.err | grep '^{' > gpurun_out/r02_v6_bench.json
timeout 600 python bench.py --workload picker --steps 8 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/r02_v6_bench_picker.json
timeout 600 python bench.py --format csvo --steps 20 --warmup 5 --skip-cpu 2>/dev/null | grep '^{' > gpurun_out/r02_v6_bench_csvo.json
timeout 600 python bench.py --workload picker --format csvo --steps 5 --warmup 2 --skip-cpu 2>/dev/null | grep '^{' > gpurun_out/r02_v6_bench_picker_csvo.json
timeout 600 python bench.py --width 7680 --height 4320 --steps 10 --warmup 3 --skip-cpu 2>/dev/null | grep '^{' > gpurun_out/r02_v6_bench_8k.json
timeout 600 python bench.py --width 1920 --height 1080 --no-shadows --steps 30 --warmup 5 --skip-cpu 2>/dev/null | grep '^{' > gpurun_out/r02_v6_bench_1080p.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/r02_v6_bench_reference.json
timeout 600 python bench.py --workload serialize --steps 20 --warmup 5 2>/dev/null | grep '^{' > gpurun_out/r02_v6_bench_serialize.json
for f in gpurun_out/r02_v6_bench*.json; do python -c "
import json,sys
d=json.load(open('$f')); e=d.get('e2e') or {}
print('$f', round(d['value'],1), d['unit'], round(d['ms_per_step'],4), 'e2e', round(e.get('value',0),1), 'frac', (d.get('roofline') or {}).get('frac'), 'issue', ((d.get('roofline') or {}).get('issue') or {}).get('frac'))"; done
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
