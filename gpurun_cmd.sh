cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/t_all_gpu.log 2>&1
tail -n 4 gpurun_out/t_all_gpu.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
tail -n 3 gpurun_out/smoke.log
timeout -s KILL 600 python bench.py > gpurun_out/bench36.json 2> gpurun_out/bench36.err
cut -c1-260 gpurun_out/bench36.json; tail -n 3 gpurun_out/bench36.err
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench36_ref.json 2> gpurun_out/bench36_ref.err
cut -c1-200 gpurun_out/bench36_ref.json
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_train36.csv python tools/profile_step.py --batch 512 > gpurun_out/prof36.log 2>&1
tail -n 1 gpurun_out/prof36.log
timeout 600 python tools/gemm_bench.py > gpurun_out/gemm_bench36.txt 2>&1
timeout 300 python tools/gemm_shapes.py > gpurun_out/gemm_shapes36.txt 2>&1
timeout 300 python tools/gn_bench.py > gpurun_out/gn_bench36.txt 2>&1
head -n 1 gpurun_out/gemm_shapes36.txt
GP_CASE=fwd256 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 2 -c 1 -f -o gpurun_out/r36_gemm_fwd256 python tools/gemm_probe.py > gpurun_out/ncu36a.log 2>&1
GP_CASE=wgrad128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 2 -c 1 -f -o gpurun_out/r36_gemm_wgrad128 python tools/gemm_probe.py > gpurun_out/ncu36b.log 2>&1
tail -n 1 gpurun_out/ncu36a.log gpurun_out/ncu36b.log
