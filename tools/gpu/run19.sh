cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
VX_BENCH_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29608 bench.py --gpus 8 --steps 20 --warmup 5 --skip-cpu 2>gpurun_out/r2_trace8.err | grep '^{' > gpurun_out/r2_trace8.json
grep trace gpurun_out/r2_trace8.err
python -c "
import json
d=json.load(open('gpurun_out/r2_trace8.json')); e=d['e2e']
print('T8', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(e['value'],1), round(e['ms_per_step'],4))"
