# round 2, session 2, run 1: GPU suite + A/Bs of the LIFO hand-over, the ray binning, two trace-kernel builds
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
fmt='
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); r=d.get("roofline",{})
        print(TAG, round(d["value"],1), round(d["ms_per_step"],4), "kernel_ms", round(r.get("kernel_ms",0),4), r.get("kernel_ms_split"), "e2e", round((d.get("e2e") or {}).get("value") or 0,1))
'
for f in "" "--no-lifo" "" "--no-lifo"; do
  timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-e2e $f 2>/dev/null | grep '^{' | tee -a gpurun_out/r3_frame_ab.jsonl | python -c "TAG='FRAME [$f]'$fmt"
done
for f in "--width 7680 --height 4320" "--width 7680 --height 4320 --no-lifo" "--format csvo" "--format csvo --no-lifo"; do
  timeout 300 python bench.py --steps 15 --warmup 4 --skip-cpu --skip-e2e $f 2>/dev/null | grep '^{' | tee -a gpurun_out/r3_frame_ab.jsonl | python -c "TAG='FRAME [$f]'$fmt"
done
for b in -1 0 24 23 22 8 7 21; do
  timeout 400 python bench.py --workload picker --steps 6 --warmup 2 --skip-cpu --skip-e2e --bin $b 2>/dev/null | grep '^{' | tee -a gpurun_out/r3_picker_ab.jsonl | python -c "TAG='PICKER [--bin $b]'$fmt"
done
for r in 8 12 16 24; do
  timeout 400 python bench.py --workload picker --steps 6 --warmup 2 --skip-cpu --skip-e2e --refill $r 2>/dev/null | grep '^{' | tee -a gpurun_out/r3_picker_ab.jsonl | python -c "TAG='PICKER [--refill $r]'$fmt"
done
echo "== variants"; timeout 600 python tools/ab_kernels.py --out gpurun_out/r3_ab1.jsonl --variants minb9,unroll2 2>&1 | tail -8
# DRAM traffic of the frame kernels, LIFO on / off (metrics only; times under ncu are not bench values)
for f in "" "--no-lifo"; do
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none \
    -k regex:"trace_primary_kernel|shade_kernel|trace_shadow_kernel" -s 5 -c 3 --csv --log-file "gpurun_out/r3_dram${f// /_}.csv" \
    python bench.py --steps 1 --warmup 3 --skip-cpu --skip-e2e $f > /dev/null 2>&1
  echo "DRAM [$f]"; python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open("gpurun_out/r3_dram${f// /_}.csv") if l.startswith('"'))]
h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value")
for r in rows[1:]: print("  ", r[ki][:40], r[mi], r[vi])
PY
done
