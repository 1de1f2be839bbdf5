# round 2, session 2, run 13 (8 GPUs): N = 8 lines (4K, 8K) with one band per frame in the e2e loop
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() {
  VX_BENCH_TRACE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 --steps 20 --warmup 5 --skip-cpu $3 2>gpurun_out/r3_n$1.err | grep '^{' | tee -a gpurun_out/r3_scale8c.jsonl | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']
print('N$1 [$3]', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(e['value'],1), round(e['ms_per_step'],4), e.get('parity_check'))"
  grep -i "trace\|error\|gave up\|Traceback" gpurun_out/r3_n$1.err | head -4
}
run 8 29661 ""
run 8 29662 "--width 7680 --height 4320"
