# ncu evidence of round 2 (1 GPU): launch list of the bench command + --set full of the frame kernels and the picker kernel.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
set -x
# 1. launch list of the bench command itself (cold-cache, serialised times: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 8 --warmup 3 --skip-cpu > gpurun_out/r02_launches_bench.log 2>&1
# 2. --set full of the three frame kernels (non-counting instances: the first 5 launches of these names are the two counting frames)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"trace_primary_kernel|shade_kernel|trace_shadow_kernel" -s 5 -c 3 \
  -f -o gpurun_out/prof_r02_frame python bench.py --steps 1 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/ncu_r02_frame.log 2>&1
# 3. the picker kernel, 16 Mi rays on the r=40 no-LOD world (first launch of the name = the counting instance)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"trace_picker_kernel" -s 1 -c 1 \
  -f -o gpurun_out/prof_r02_picker python bench.py --workload picker --steps 1 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/ncu_r02_picker.log 2>&1
# 4. the frame kernels on the CSVO world
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"trace_primary_kernel|shade_kernel|trace_shadow_kernel" -s 5 -c 3 \
  -f -o gpurun_out/prof_r02_csvo python bench.py --format csvo --steps 1 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/ncu_r02_csvo.log 2>&1
ls -la gpurun_out/*.ncu-rep
# bench lines of this build (not under a profiler)
timeout 600 python bench.py --steps 30 --warmup 5 2>gpurun_out/r02_v5_bench.err | grep '^{' > gpurun_out/r02_v5_bench.json
timeout 600 python bench.py --workload picker --steps 8 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/r02_v5_bench_picker.json
timeout 600 python bench.py --format csvo --steps 20 --warmup 5 --skip-cpu 2>/dev/null | grep '^{' > gpurun_out/r02_v5_bench_csvo.json
timeout 600 python bench.py --workload picker --format csvo --steps 5 --warmup 2 --skip-cpu 2>/dev/null | grep '^{' > gpurun_out/r02_v5_bench_picker_csvo.json
timeout 600 python bench.py --width 7680 --height 4320 --steps 10 --warmup 3 --skip-cpu 2>/dev/null | grep '^{' > gpurun_out/r02_v5_bench_8k.json
timeout 600 python bench.py --width 1920 --height 1080 --no-shadows --steps 30 --warmup 5 --skip-cpu 2>/dev/null | grep '^{' > gpurun_out/r02_v5_bench_1080p.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/r02_v5_bench_reference.json
timeout 600 python bench.py --workload serialize --steps 20 --warmup 5 2>/dev/null | grep '^{' > gpurun_out/r02_v5_bench_serialize.json
for f in gpurun_out/r02_v5_bench*.json; do python -c "
import json,sys
d=json.load(open('$f')); e=d.get('e2e') or {}
print('$f', round(d['value'],1), d['unit'], round(d['ms_per_step'],4), 'e2e', round(e.get('value',0),1))"; done
