cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t54_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/t54_gpu.log
timeout 400 python bench.py > gpurun_out/bench54.json 2> gpurun_out/bench54.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench54.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches54.csv python tools/profile_step.py --batch 512 > gpurun_out/ncu54_list.log 2>&1; echo "ncu list rc=$?"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke54.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke54.log
