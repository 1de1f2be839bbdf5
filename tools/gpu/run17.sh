cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --skip-cpu 2>gpurun_out/r2_n2b.err | grep '^{' > gpurun_out/r2_n2b.json; tail -3 gpurun_out/r2_n2b.err
python -c "
import json
d=json.load(open('gpurun_out/r2_n2b.json')); e=d['e2e']
print('N2', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(e['value'],1), round(e['ms_per_step'],4), e.get('parity_check'))"
timeout 600 python bench.py --gpus 2 --group --steps 20 --warmup 5 2>gpurun_out/r2_g2b.err | grep '^{' > gpurun_out/r2_g2b.json; tail -3 gpurun_out/r2_g2b.err
python -c "
import json
d=json.load(open('gpurun_out/r2_g2b.json')); e=d['e2e']
print('G2', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(e['value'],1), round(e['ms_per_step'],4))"
