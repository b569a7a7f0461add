# 8-GPU run: data-parallel bench (our arm) + 4 GPUs
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; echo "8gpu rc=$?"; cut -c1-300 gpurun_out/r02_bench_8gpu.json; tail -3 gpurun_out/r02_bench_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02_bench_4gpu.json 2> gpurun_out/r02_bench_4gpu.err; echo "4gpu rc=$?"; cut -c1-300 gpurun_out/r02_bench_4gpu.json
timeout 300 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/r02_bench_1of8.json 2> gpurun_out/r02_bench_1of8.err; echo "1gpu rc=$?"; cut -c1-300 gpurun_out/r02_bench_1of8.json
