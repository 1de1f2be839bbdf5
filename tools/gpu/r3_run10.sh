# round 2, session 2, run 10 (2 GPUs): shade CTA = 4 strips (GPU suite + A/B of 1/2/8), N=2 e2e with one band per frame
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== strips per shade CTA"; timeout 600 python tools/ab_kernels.py --out gpurun_out/r3_ab4.jsonl --variants shade1,shade2,shade8 2>&1 | tail -5
VX_BENCH_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 20 --warmup 5 --skip-cpu 2>gpurun_out/r3_n2.err | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']
print('N2 [1 band]', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(e['value'],1), round(e['ms_per_step'],4), e.get('parity_check'))"
grep -i "trace\|error\|gave up\|Traceback" gpurun_out/r3_n2.err | head
VX_BENCH_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 2 --steps 20 --warmup 5 --skip-cpu --bands-pipelined 2 2>gpurun_out/r3_n2.err | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']
print('N2 [2 bands]', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(e['value'],1), round(e['ms_per_step'],4), e.get('parity_check'))"
