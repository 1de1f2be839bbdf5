# round 2, session 2, run 18: what one rank of an 8 / 4-GPU frame does, on ONE GPU (--sim-shard): CTAs per SM of the persistent kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
fmt='
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); r=d.get("roofline",{})
        print(TAG, "step", round(d["ms_per_step"],4), "kernels", round(r.get("kernel_ms",0),4), r.get("kernel_ms_split"))
'
for s in 8 4 2; do
for c in 0 7 6 5 4; do
  timeout 200 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-e2e --sim-shard $s --ctas-per-sm $c 2>/dev/null | grep '^{' | python -c "TAG='SIM$s [ctas $c]'$fmt"
done
done
