cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t55_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/t55_gpu.log
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --sample-steps 8 > gpurun_out/bench55.json 2> gpurun_out/bench55.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench55.json
