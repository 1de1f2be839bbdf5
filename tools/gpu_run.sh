# usage: bash tools/gpu_run.sh [test] [sweep "<opts>;<opts>;..."] [ncu <name> "<bench opts>"] [launches <name>]
# Everything it writes goes to gpurun_out/.
fmt='
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); r=d.get("roofline",{})
        print(round(d["value"]),d["unit"],round(d["ms_per_step"],3),"ms/step kernel_ms",round(r.get("kernel_ms",0),3), r.get("kernel_ms_split"), "frac",round(r.get("frac",0),4), "e2e", round((d.get("e2e") or {}).get("value") or 0), round((d.get("e2e") or {}).get("ms_per_step") or 0,3), "issue_ms", d.get("host_issue_ms_per_step"))
    else: print(l)
'
while [ $# -gt 0 ]; do
  case "$1" in
    test) python -m pytest tests -m gpu -x -q 2>&1 | tail -15; shift;;
    sweep) IFS=';' read -ra OPTS <<< "$2"; for opt in "${OPTS[@]}"; do echo "== $opt"; python bench.py --steps 10 --warmup 3 --skip-cpu --skip-e2e $opt 2>&1 | python -c "$fmt"; done; shift 2;;
    sweepe) IFS=';' read -ra OPTS <<< "$2"; for opt in "${OPTS[@]}"; do echo "== $opt"; python bench.py --steps 10 --warmup 3 --skip-cpu $opt 2>&1 | python -c "$fmt"; done; shift 2;;
    ncu) ncu --set full --clock-control none --import-source on -k regex:"trace_|shade_" -s 6 -c 3 -f -o gpurun_out/prof_$2 python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e $3 > gpurun_out/ncu_$2.log 2>&1; tail -3 gpurun_out/ncu_$2.log; shift 3;;
    launches) ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$2.csv python bench.py --steps 8 --warmup 3 --skip-cpu > gpurun_out/launches_$2.log 2>&1; tail -2 gpurun_out/launches_$2.log | cut -c1-300; shift 2;;
    sanitize) compute-sanitizer --tool $2 --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or edge or pipelined or sharded_frames or dirty" > gpurun_out/sanitizer_$2.log 2>&1; echo "sanitizer $2 rc=$?"; tail -5 gpurun_out/sanitizer_$2.log; shift 2;;
    mtest) python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -15; shift;;
    scale1) echo "== N=$2 $3"; python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $2 --steps 20 --warmup 5 --skip-cpu $3 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|^$\|NCCL version" | tee -a gpurun_out/scale_lines.jsonl | python -c "$fmt"; shift 3;;
    scale) for opt in "" "--gather nccl"; do echo "== N=$2 $opt $3"; python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $2 --steps 20 --warmup 5 --skip-cpu $opt $3 2>&1 | grep -v "^W\|^\*\*\*" | python -c "$fmt"; done; shift 3;;
    ncupick) ncu --set full --clock-control none --import-source on -k regex:"trace_picker" -s 2 -c 1 -f -o gpurun_out/prof_$2 python bench.py --workload picker --steps 1 --warmup 1 --skip-cpu --skip-e2e $3 > gpurun_out/ncu_$2.log 2>&1; tail -3 gpurun_out/ncu_$2.log | cut -c1-400; shift 3;;
    bench) python bench.py $2 > gpurun_out/bench_$3.json 2> gpurun_out/bench_$3.err; cat gpurun_out/bench_$3.json | python -c "$fmt"; shift 3;;
    *) echo "unknown $1"; shift;;
  esac
done
