# round 2, session 2, run 15: picker kernel without the unorm table (9 resident CTAs per SM instead of 8)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "picker or edge or golden or mc_world" 2>&1 | tail -3
for f in "" "--format csvo"; do
timeout 400 python bench.py --workload picker --steps 8 --warmup 3 --skip-cpu $f 2>/dev/null | grep '^{' | tee gpurun_out/r3_picker9$(echo $f | tr -d ' -').json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('PICKER [$f]', round(d['value'],1), round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value'],1))"
done
