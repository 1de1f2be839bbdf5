# round 2, session 2, run 12 (4 GPUs): GPU-side phase marks of the N > 1 e2e loop
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
VX_BENCH_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 4 --steps 20 --warmup 5 --skip-cpu 2>gpurun_out/r3_n4.err | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']
print('N4', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(e['value'],1), round(e['ms_per_step'],4), e.get('parity_check'))"
grep -i "trace\|error\|gave up\|Traceback" gpurun_out/r3_n4.err | head
