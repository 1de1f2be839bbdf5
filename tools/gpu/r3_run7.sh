# round 2, session 2, run 7: shadow list grouped by face inside a CTA's segment (variant build): parity + timing; picker refill re-tune
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
LIVE=voxel-rs_b200/libvoxelrt.so
cp $LIVE /tmp/live.so
cp voxel-rs_b200/variants/shadow_by_face/libvoxelrt.so $LIVE; touch $LIVE
echo "== parity on the variant"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "terrain_variants or full_size_4k or reference_scene or sharded_frames or pipelined" 2>&1 | tail -4
cp /tmp/live.so $LIVE; touch $LIVE
echo "== timing"; timeout 600 python tools/ab_kernels.py --out gpurun_out/r3_ab3.jsonl --variants shadow_by_face 2>&1 | tail -4
fmt='
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); r=d.get("roofline",{}); e=d.get("e2e") or {}
        print(TAG, round(d["value"],1), round(d["ms_per_step"],4), "kernel_ms", round(r.get("kernel_ms",0),4))
'
for r in 12 16 20 24 28; do
  timeout 400 python bench.py --workload picker --steps 6 --warmup 2 --skip-cpu --skip-e2e --refill $r 2>/dev/null | grep '^{' | python -c "TAG='PICKER [--refill $r]'$fmt"
done
