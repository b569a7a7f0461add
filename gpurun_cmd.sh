# GPU run 30 (one B200): FIR tap loop with packed fp32 FMAs (FFMA2)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
for v in 0 1; do
echo "== ST_FIR_FFMA2=$v"; ST_FIR_FFMA2=$v timeout 150 python tools/upfirdn_bench.py 2>&1 | grep -v "^st_upfirdn2d" | cut -c1-130
done > gpurun_out/fir_ffma2.txt 2>&1
cat gpurun_out/fir_ffma2.txt
ST_FIR_FFMA2=1 timeout 200 python -m pytest tests/test_gpu_round2.py tests/test_gpu_kernels.py tests/test_gpu_parity.py -q -k "upfirdn or fir or full_width" --timeout=150 2>&1 | tail -1
