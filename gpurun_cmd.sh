cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 600 -k "langevin or predictors or mixed" > gpurun_out/t5_new.log 2>&1
tail -n 30 gpurun_out/t5_new.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_train18.csv python tools/profile_step.py --batch 512 > gpurun_out/prof18.log 2>&1
tail -n 2 gpurun_out/prof18.log
GP_CASE=fwd256 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 2 -c 1 -f -o gpurun_out/r18_gemm_fwd256 python tools/gemm_probe.py > gpurun_out/ncu18a.log 2>&1
GP_CASE=wgrad128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 2 -c 1 -f -o gpurun_out/r18_gemm_wgrad128 python tools/gemm_probe.py > gpurun_out/ncu18b.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gn_ -s 40 -c 5 -f -o gpurun_out/r18_gn python tools/profile_step.py --batch 512 > gpurun_out/ncu18c.log 2>&1
tail -n 2 gpurun_out/ncu18a.log gpurun_out/ncu18b.log gpurun_out/ncu18c.log
ls -la gpurun_out/*.ncu-rep
