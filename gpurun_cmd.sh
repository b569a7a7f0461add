# GPU run 20 (one B200): FIR kernel with batched footprint loads
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_kernels.py tests/test_gpu_parity.py -q -k "upfirdn or fir or full_width" --timeout=300 > gpurun_out/t_fir.log 2>&1; echo "fir tests rc=$?"; tail -n 3 gpurun_out/t_fir.log
timeout 300 python tools/upfirdn_bench.py > gpurun_out/r02_upfirdn_bench.txt 2>&1; echo "upfirdn rc=$?"; cat gpurun_out/r02_upfirdn_bench.txt
