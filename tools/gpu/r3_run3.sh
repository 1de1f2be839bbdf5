# round 2, session 2, run 3: GPU suite with the two-frames-in-flight read-back; e2e loop variants
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
fmt='
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); r=d.get("roofline",{}); e=d.get("e2e") or {}
        print(TAG, round(d["value"],1), round(d["ms_per_step"],4), "kernel_ms", round(r.get("kernel_ms",0),4), "e2e", round(e.get("value") or 0,1), round(e.get("ms_per_step") or 0,4), "blocking", round((e.get("blocking_call") or {}).get("value") or 0,1), round((e.get("blocking_call") or {}).get("ms_per_step") or 0,4), "issue", d.get("host_issue_ms_per_step"))
'
for f in "" "--bands-pipelined 2 --bands 2" "--bands-pipelined 3 --bands 4" "--bands 5" ""; do
  timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu $f 2>gpurun_out/r3_e2e.err | grep '^{' | tee -a gpurun_out/r3_e2e_ab.jsonl | python -c "TAG='E2E [$f]'$fmt"
  tail -3 gpurun_out/r3_e2e.err
done
for f in "--width 7680 --height 4320" "--width 1920 --height 1080 --no-shadows" "--format csvo"; do
  timeout 300 python bench.py --steps 15 --warmup 4 --skip-cpu $f 2>/dev/null | grep '^{' | tee -a gpurun_out/r3_e2e_ab.jsonl | python -c "TAG='E2E [$f]'$fmt"
done
