cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() {
  n=$1; tag=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 20 --warmup 5 --skip-cpu "$@" 2>gpurun_out/r2_scale2_$tag.err | grep '^{' > gpurun_out/r2_scale2_$tag.json
  python -c "
import json
d=json.load(open('gpurun_out/r2_scale2_$tag.json')); e=d.get('e2e',{})
print('SCALE2 $tag', d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(e.get('value',0),1), round(e.get('ms_per_step',0),4), e.get('parity_check','')[:30])" || tail -5 gpurun_out/r2_scale2_$tag.err
}
run 8 4k_n8
run 4 4k_n4
run 8 8k_n8 --width 7680 --height 4320
timeout 600 python bench.py --gpus 8 --group --steps 20 --warmup 5 2>gpurun_out/r2_group2_n8.err | grep '^{' > gpurun_out/r2_group2_n8.json
python -c "
import json
d=json.load(open('gpurun_out/r2_group2_n8.json')); e=d['e2e']
print('GROUP2', d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(e['value'],1), round(e['ms_per_step'],4))" || tail -5 gpurun_out/r2_group2_n8.err
