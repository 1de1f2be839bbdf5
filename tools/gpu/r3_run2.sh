# round 2, session 2, run 2: where the binned picker batch loses its time; LIFO hand-over split into its parts
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
fmt='
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); r=d.get("roofline",{})
        print(TAG, round(d["value"],1), round(d["ms_per_step"],4), "kernel_ms", round(r.get("kernel_ms",0),4), r.get("kernel_ms_split"), "e2e", round((d.get("e2e") or {}).get("value") or 0,1))
'
for b in 23 0; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r3_picker_launches_bin$b.csv \
    python bench.py --workload picker --steps 2 --warmup 1 --skip-cpu --skip-e2e --bin $b > /dev/null 2>&1
  echo "PICKER LAUNCHES bin=$b"; python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open("gpurun_out/r3_picker_launches_bin$b.csv") if l.startswith('"'))]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
for r in rows[-14:]: print("  ", r[ki][:60], r[vi], r[ui])
PY
done
for l in 0 1 3 7 0 1 3 7; do
  timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-e2e --lifo $l 2>/dev/null | grep '^{' | tee -a gpurun_out/r3_lifo_ab.jsonl | python -c "TAG='FRAME [--lifo $l]'$fmt"
done
for l in 0 3 7; do
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --cache-control none --clock-control none \
    -k regex:"trace_primary_kernel|shade_kernel|trace_shadow_kernel" -s 5 -c 3 --csv --log-file "gpurun_out/r3_dram_lifo$l.csv" \
    python bench.py --steps 1 --warmup 3 --skip-cpu --skip-e2e --lifo $l > /dev/null 2>&1
  echo "DRAM [--lifo $l] (cache-control none)"; python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open("gpurun_out/r3_dram_lifo$l.csv") if l.startswith('"'))]
h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value")
for r in rows[1:]: print("  ", r[ki][:40], r[mi], r[vi])
PY
done
