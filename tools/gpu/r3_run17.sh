# round 2, session 2, run 17: compute-sanitizer (memcheck, racecheck) over the GPU tests that exercise this round's new device code:
# LIFO discards, ray binning kernels, the 4-strip shade loop, two frames in flight, scatter / bounds kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
K="golden or edge or pipelined or sharded_frames or dirty or terrain_variants or primary_hits"
for tool in memcheck racecheck; do
  timeout 700 compute-sanitizer --tool $tool --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" > gpurun_out/r02_sanitizer_$tool.txt 2>&1
  echo "sanitizer $tool rc=$?"; tail -4 gpurun_out/r02_sanitizer_$tool.txt
done
# the binning kernels: one small scene of the 200k-ray test is enough (atomics + scans + scatter)
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python - > gpurun_out/r02_sanitizer_binning.txt 2>&1 <<'PY'
import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import __graft_entry__ as graft, helpers
graft.build(); pkg = graft.load_pkg(); ora = graft.load_oracle()
reg = pkg.content_registry(pkg.load_atlas())
from test_gpu_parity import small_scenes, make_svo
name, blocks, svo_pos = next(iter(small_scenes(pkg)))
w = helpers.shader_test_world(pkg, blocks, svo_pos)
s = helpers.oracle_scene(ora, w, reg)
svo = make_svo(pkg, reg, w, size_mb=4, w=8, h=8, rays=1 << 18)
tasks = helpers.random_tasks(pkg, 100_000, -8.0, 32 * (max(svo_pos) + 1) + 8.0, -1.0, 1)
want, _ = s.raycast(tasks)
for b in (7 | 16, 8, 3):
    svo.set_option(15, b)
    got = svo.raycast_tasks(tasks)
    assert got.tobytes() == want.tobytes(), b
print("binning under memcheck: 3 settings x 100k rays byte-identical to the oracle")
PY
echo "sanitizer binning rc=$?"; tail -4 gpurun_out/r02_sanitizer_binning.txt
