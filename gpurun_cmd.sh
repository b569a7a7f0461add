# GPU run 19 (one B200): full GPU suite, smoke, default bench line (with the reference legs), reference arm
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 1500 python -m pytest tests -q -m gpu --timeout=900 > gpurun_out/t_gpu_final.log 2>&1; echo "pytest -m gpu rc=$?"; tail -n 4 gpurun_out/t_gpu_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke_final.log
timeout 900 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02_bench_1gpu.json; tail -3 gpurun_out/r02_bench_1gpu.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "reference arm rc=$?"; cut -c1-300 gpurun_out/r02_bench_reference_arm.json
