# round 2, session 2, run 21: slice size of the pipelined vx_raycast (upload / trace / read-back overlap), picker e2e
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
LIVE=voxel-rs_b200/libvoxelrt.so
cp $LIVE /tmp/live.so
for v in default slice19 slice20 slice22; do
  if [ $v = default ]; then cp /tmp/live.so $LIVE; else cp voxel-rs_b200/variants/$v/libvoxelrt.so $LIVE; fi; touch $LIVE
  timeout 300 python bench.py --workload picker --steps 4 --warmup 2 --skip-cpu 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']; print('PICKER e2e $v', round(d['value'],1), 'e2e', round(e['value'],1), round(e['ms_per_step'],3))"
done
cp /tmp/live.so $LIVE; touch $LIVE
