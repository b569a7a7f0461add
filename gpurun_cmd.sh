cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t47_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/t47_gpu.log
timeout 400 python bench.py > gpurun_out/bench47.json 2> gpurun_out/bench47.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/bench47.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench47_ref.json 2> gpurun_out/bench47_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench47_ref.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches47.csv python tools/profile_step.py --batch 512 > gpurun_out/ncu47_list.log 2>&1; echo "ncu list rc=$?"
timeout 300 python tools/gn_bench.py > gpurun_out/gn_bench47.txt 2>&1; echo "gnbench rc=$?"
timeout 300 python tools/gemm_shapes.py > gpurun_out/gemm_shapes47.txt 2>&1; echo "gemm_shapes rc=$?"; head -n 2 gpurun_out/gemm_shapes47.txt
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gn_fwd_fused_kernel -c 3 -f -o gpurun_out/ncu47_gnfwd python tools/profile_step.py --batch 512 > gpurun_out/ncu47_gnfwd.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*47*
