# round 2, session 2, run 19 (2 GPUs): uploads on the library's own stream (scatter + bounds next to the L2 flush) vs everything on one stream
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
fmt='
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); e=d.get("e2e") or {}
        print(TAG, round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(e.get("value") or 0,1), round(e.get("ms_per_step") or 0,4), e.get("parity_check"), (d["config"].get("parity_check") or "")[:30])
'
for f in "" "--one-stream"; do
  timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu $f 2>/dev/null | grep '^{' | python -c "TAG='N1 [$f]'$fmt"
done
p=29671
for f in "" "--one-stream" ""; do
  p=$((p+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $p bench.py --gpus 2 --steps 20 --warmup 5 --skip-cpu $f 2>gpurun_out/r3_n2.err | grep '^{' | python -c "TAG='N2 [$f]'$fmt"
  grep -i "error\|gave up\|Traceback" gpurun_out/r3_n2.err | head -3
done
