# round 2, session 2, run 4: steps-per-vote variants of the walk loops; D2H bandwidth in the e2e line
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
fmt='
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); r=d.get("roofline",{}); e=d.get("e2e") or {}
        print(TAG, round(d["value"],1), round(d["ms_per_step"],4), "kernel_ms", round(r.get("kernel_ms",0),4), r.get("kernel_ms_split"), "e2e", round(e.get("value") or 0,1), e.get("pcie_d2h"))
'
echo "== frame variants"; timeout 600 python tools/ab_kernels.py --out gpurun_out/r3_ab2.jsonl --variants prim_seq2,prim_one,shadow_seq2 2>&1 | tail -8
LIVE=voxel-rs_b200/libvoxelrt.so
cp $LIVE /tmp/live.so
for v in csvo_two default; do
  if [ $v = default ]; then cp /tmp/live.so $LIVE; else cp voxel-rs_b200/variants/$v/libvoxelrt.so $LIVE; fi; touch $LIVE
  timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-e2e --format csvo 2>/dev/null | grep '^{' | python -c "TAG='CSVO [$v]'$fmt"
done
for v in thresh_two default; do
  if [ $v = default ]; then cp /tmp/live.so $LIVE; else cp voxel-rs_b200/variants/$v/libvoxelrt.so $LIVE; fi; touch $LIVE
  timeout 400 python bench.py --workload picker --steps 6 --warmup 2 --skip-cpu --skip-e2e 2>/dev/null | grep '^{' | python -c "TAG='PICKER [$v]'$fmt"
  timeout 400 python bench.py --workload picker --steps 6 --warmup 2 --skip-cpu --skip-e2e --format csvo 2>/dev/null | grep '^{' | python -c "TAG='PICKER CSVO [$v]'$fmt"
done
cp /tmp/live.so $LIVE; touch $LIVE
timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu 2>/dev/null | grep '^{' | python -c "TAG='E2E []'$fmt"
