cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
run() {  # n, extra args, tag
  n=$1; tag=$2; shift 2
  if [ $n -eq 1 ]; then timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --skip-cpu "$@" 2>gpurun_out/r2_scale_$tag.err | grep '^{' > gpurun_out/r2_scale_$tag.json
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 20 --warmup 5 --skip-cpu "$@" 2>gpurun_out/r2_scale_$tag.err | grep '^{' > gpurun_out/r2_scale_$tag.json; fi
  python -c "
import json
d=json.load(open('gpurun_out/r2_scale_$tag.json')); e=d.get('e2e',{})
print('SCALE $tag', d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(e.get('value',0),1), round(e.get('ms_per_step',0),4), d['roofline']['kernel_ms_split'], d['config'].get('parity_check','')[:40], e.get('parity_check','')[:30])" || tail -5 gpurun_out/r2_scale_$tag.err
}
for n in 1 2 4 8; do run $n 4k_n$n; done
for n in 1 2 4 8; do run $n 8k_n$n --width 7680 --height 4320; done
for n in 2 8; do
  timeout 600 python bench.py --gpus $n --group --steps 20 --warmup 5 2>gpurun_out/r2_group_n$n.err | grep '^{' > gpurun_out/r2_group_n$n.json
  python -c "
import json
d=json.load(open('gpurun_out/r2_group_n$n.json')); e=d['e2e']
print('GROUP', d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(e['value'],1), round(e['ms_per_step'],4))" || tail -5 gpurun_out/r2_group_n$n.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 8 --workload picker --steps 5 --warmup 2 --skip-cpu 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('PICK8', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))"
