cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "groupnorm or colsum" > gpurun_out/t40_gn.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/t40_gn.log
timeout 300 python tools/gn_bench.py > gpurun_out/gn_bench40.txt 2>&1; echo "gnbench rc=$?"; cat gpurun_out/gn_bench40.txt
for f in 0 1; do
ST_GN_FUSED=$f timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --sample-steps 4 > gpurun_out/bench40_f$f.json 2> gpurun_out/bench40_f$f.err; echo "bench f=$f rc=$?"; cut -c1-220 gpurun_out/bench40_f$f.json
done
