# GPU run 29 (one B200): final full GPU suite + smoke on the round's last build
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/test_stats.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests -q -m gpu --timeout=600 > gpurun_out/t_gpu_final.log 2>&1; echo "pytest -m gpu rc=$?"; tail -n 2 gpurun_out/t_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke_final.log
