# round 2, session 2, run 8 (2 GPUs): the N > 1 e2e loop with two frames in flight — parity + phase trace; sharded / group GPU tests
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== multi-GPU tests"; timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_group.py -m gpu -x -q 2>&1 | tail -4
for f in "" "--width 7680 --height 4320"; do
VX_BENCH_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 --skip-cpu $f 2>gpurun_out/r3_n2.err | grep '^{' | tee -a gpurun_out/r3_scale2.jsonl | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']
print('N2 [$f]', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(e['value'],1), round(e['ms_per_step'],4), e.get('parity_check'), d['config'].get('parity_check','')[:60])"
grep -i "trace\|error\|gave up\|Traceback" gpurun_out/r3_n2.err | head
done
