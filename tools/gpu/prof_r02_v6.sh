# ncu evidence of the round-2 final build (1 GPU): launch list of the bench command + --set full of the frame kernels (ESVO, CSVO) and the picker kernel.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
set -x
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_v6_launches.csv \
  python bench.py --steps 8 --warmup 3 --skip-cpu > gpurun_out/r02_v6_launches_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"trace_primary_kernel|shade_kernel|trace_shadow_kernel" -s 5 -c 3 \
  -f -o gpurun_out/prof_r02v6_frame python bench.py --steps 1 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/ncu_r02v6_frame.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"trace_picker_kernel" -s 1 -c 1 \
  -f -o gpurun_out/prof_r02v6_picker python bench.py --workload picker --steps 1 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/ncu_r02v6_picker.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"trace_primary_kernel|shade_kernel|trace_shadow_kernel" -s 5 -c 3 \
  -f -o gpurun_out/prof_r02v6_csvo python bench.py --format csvo --steps 1 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/ncu_r02v6_csvo.log 2>&1
ls -la gpurun_out/prof_r02v6*.ncu-rep
tail -3 gpurun_out/ncu_r02v6_frame.log | cut -c1-300
