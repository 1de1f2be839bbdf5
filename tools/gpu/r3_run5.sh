# round 2, session 2, run 5: where the pipelined e2e loop spends 0.2 ms more than the resident loop
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
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
timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --no-flush 2>/dev/null | grep '^{' | python -c "TAG='E2E no-flush'$fmt"
VX_BENCH_SKIP_COMMIT=1 timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --no-flush 2>/dev/null | grep '^{' | python -c "TAG='E2E no-commit no-flush'$fmt"
timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --bands-pipelined 2 2>/dev/null | grep '^{' | python -c "TAG='E2E bands-pipelined 2'$fmt"
