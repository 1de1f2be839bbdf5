set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_group.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_n2_bench.json 2> gpurun_out/r2_n2_bench.err; tail -c 1500 gpurun_out/r2_n2_bench.json; tail -5 gpurun_out/r2_n2_bench.err
timeout 600 python bench.py --gpus 2 --group --steps 20 --warmup 5 > gpurun_out/r2_n2_group.json 2> gpurun_out/r2_n2_group.err; tail -c 1200 gpurun_out/r2_n2_group.json; tail -5 gpurun_out/r2_n2_group.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2_n2_reference.json 2> gpurun_out/r2_n2_reference.err; cut -c1-900 gpurun_out/r2_n2_reference.json
