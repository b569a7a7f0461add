# GPU run 7 (one B200): halo form of the 3x3 convolution
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/test_stats.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 600 python -m pytest tests/test_gpu_round2.py -q -k "halo or bench_shapes" --timeout=300 > gpurun_out/t_halo.log 2>&1; HL=$?; echo "halo check rc=$HL"; tail -n 12 gpurun_out/t_halo.log; grep "halo vs" gpurun_out/test_stats.txt | head -20
if [ $HL -ne 0 ]; then export ST_TC_HALO=0; echo "falling back to ST_TC_HALO=0 for the rest"; fi
timeout 600 python tools/gemm_bench.py > gpurun_out/gemm_bench_r2f.txt 2>&1; echo "gemm_bench rc=$?"; head -34 gpurun_out/gemm_bench_r2f.txt
timeout 400 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_r2f.json
ST_TC_HALO=0 timeout 400 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_r2f_nohalo.json 2> gpurun_out/bench_r2f_nohalo.err; echo "nohalo rc=$?"; cut -c1-300 gpurun_out/bench_r2f_nohalo.json
timeout 1500 python -m pytest tests -q -m gpu --timeout=900 > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/t_gpu.log
