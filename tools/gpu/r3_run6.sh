# round 2, session 2, run 6: GPU suite + e2e after the staging DMA moved to its own stream
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
fmt='
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); r=d.get("roofline",{}); e=d.get("e2e") or {}
        print(TAG, round(d["value"],1), round(d["ms_per_step"],4), "kernel_ms", round(r.get("kernel_ms",0),4), "e2e", round(e.get("value") or 0,1), round(e.get("ms_per_step") or 0,4), "blocking", round((e.get("blocking_call") or {}).get("ms_per_step") or 0,4))
'
VX_BENCH_TRACE=1 timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu 2>gpurun_out/r3_t1.err | grep '^{' | python -c "TAG='E2E trace'$fmt"; grep trace gpurun_out/r3_t1.err
VX_BENCH_SKIP_COMMIT=1 timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu 2>/dev/null | grep '^{' | python -c "TAG='E2E no-commit'$fmt"
timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --bands 2 2>/dev/null | grep '^{' | python -c "TAG='E2E bands 2'$fmt"
timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --bands 4 2>/dev/null | grep '^{' | python -c "TAG='E2E bands 4'$fmt"
timeout 300 python bench.py --steps 15 --warmup 5 --skip-cpu --width 1920 --height 1080 --no-shadows 2>/dev/null | grep '^{' | python -c "TAG='E2E 1080p'$fmt"
