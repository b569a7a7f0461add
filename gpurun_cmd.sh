cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "likelihood or ode_sampler or deepest" > gpurun_out/t42_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/t42_gpu.log
for i in 1 2; do for f in 0 1; do
ST_GN_FUSED=$f timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --sample-steps 2 > gpurun_out/bench42_f${f}_$i.json 2> gpurun_out/bench42_f${f}_$i.err; echo "bench f=$f rc=$?"; cut -c1-200 gpurun_out/bench42_f${f}_$i.json | cut -c40-200
done; done
