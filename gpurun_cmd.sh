cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/bench49_8gpu.json 2> gpurun_out/bench49_8gpu.err; echo "rc=$?"; wc -l gpurun_out/bench49_8gpu.json; cut -c1-230 gpurun_out/bench49_8gpu.json; tail -n 3 gpurun_out/bench49_8gpu.err
