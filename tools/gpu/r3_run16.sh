# round 2, session 2, run 16: picker kernel with stacks-only shared memory: 9 CTAs/SM (56 registers) vs 10 (48 registers, spills outside the loops)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
LIVE=voxel-rs_b200/libvoxelrt.so
cp $LIVE /tmp/live.so
for v in default picker10; do
  if [ $v = default ]; then cp /tmp/live.so $LIVE; else cp voxel-rs_b200/variants/$v/libvoxelrt.so $LIVE; fi; touch $LIVE
  echo "== $v"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "picker or edge or golden or mc_world" 2>&1 | tail -2
  for f in "" "--format csvo" "--refill 24" "--refill 16"; do
  timeout 400 python bench.py --workload picker --steps 8 --warmup 3 --skip-cpu --skip-e2e $f 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('PICKER $v [$f]', round(d['value'],1), round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4))"
  done
done
cp /tmp/live.so $LIVE; touch $LIVE
