# 8-GPU run: BASELINE configs[2..4] data-parallel lines + the C5 sampler on 8 GPUs
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
P=29530
for c in c3 c4 c5; do
P=$((P+1))
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r02_bench_${c}_8gpu.json 2> gpurun_out/r02_bench_${c}_8gpu.err; echo "$c 8gpu rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/r02_bench_${c}_8gpu.json').read().strip().splitlines()[-1]);print(d['value'], d['ms_per_step'], d['n_gpus'])")"; tail -2 gpurun_out/r02_bench_${c}_8gpu.err | cut -c1-200
done
P=$((P+1))
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --config c5 --mode sampler --sample-steps 500 --no-cpu-baseline --no-gpu-reference > gpurun_out/r02_bench_sampler_c5_8gpu_n500.json 2> gpurun_out/r02_bench_sampler_c5_8gpu_n500.err; echo "c5 sampler 8gpu rc=$?"; cut -c1-260 gpurun_out/r02_bench_sampler_c5_8gpu_n500.json
P=$((P+1))
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --config c2 --mode sampler --sample-steps 300 --no-cpu-baseline --no-gpu-reference > gpurun_out/r02_bench_sampler_c2_8gpu_n300.json 2> gpurun_out/r02_bench_sampler_c2_8gpu_n300.err; echo "c2 sampler 8gpu rc=$?"; cut -c1-260 gpurun_out/r02_bench_sampler_c2_8gpu_n300.json
