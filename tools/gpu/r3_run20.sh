# round 2, session 2, run 20 (8 GPUs): the final arrangement at N = 8, 4K
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
VX_BENCH_TRACE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29681 bench.py --gpus 8 --steps 20 --warmup 5 --skip-cpu 2>gpurun_out/r3_n8.err | grep '^{' | tee -a gpurun_out/r3_scale8d.jsonl | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']
print('N8', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(e['value'],1), round(e['ms_per_step'],4), e.get('parity_check'), (d['config'].get('parity_check') or '')[:30])"
grep -i "trace\|error\|gave up\|Traceback" gpurun_out/r3_n8.err | head -6
