# GPU run 10 (one B200): branch-free FIR kernel; ncu full captures with the tensor-pipe metric
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_kernels.py tests/test_gpu_parity.py -q -k "upfirdn or fir or full_width" --timeout=300 > gpurun_out/t_fir.log 2>&1; FR=$?; echo "fir check rc=$FR"; tail -n 6 gpurun_out/t_fir.log
timeout 300 python tools/upfirdn_bench.py > gpurun_out/r02_upfirdn_bench.txt 2>&1; echo "upfirdn rc=$?"; cat gpurun_out/r02_upfirdn_bench.txt
timeout 600 python bench.py --config c3 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; echo "c3 rc=$?"; cut -c1-200 gpurun_out/r02_bench_c3.json
for spec in "gemm_fwd:gemm_tc2_kernel:3:6" "gemm_bwd:gemm_tc2_kernel:330:6" "other:attn_fwd_kernel|gn_apply_kernel|gn_bwd_resident:2:4"; do
  IFS=: read name rx skip cnt <<< "$spec"
  timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:"$rx" -s $skip -c $cnt -o gpurun_out/prof_$name python tools/profile_step.py --batch 512 > gpurun_out/ncu_full_$name.log 2>&1; echo "ncu full $name rc=$?"
  python tools/ncu_summary.py gpurun_out/prof_$name.ncu-rep > gpurun_out/r02_ncu_$name.md 2>> gpurun_out/ncu_full_$name.log
  ncu -i gpurun_out/prof_$name.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(l for l in sys.stdin if not l.startswith('==')))
hdr = rows[0]
keep = [i for i, h in enumerate(hdr) if h in ('ID', 'Kernel Name') or any(k in h for k in ('pipe_tensor', 'dram__bytes', 'gpu__time_duration', 'l1tex__data_pipe', 'smsp__inst_executed.sum', 'lts__t_sector_hit_rate', 'sm__throughput', 'dram__throughput', 'l1tex__data_bank', 'shared'))]
w = csv.writer(sys.stdout)
for r in rows:
  w.writerow([r[i] for i in keep])
" > gpurun_out/r02_ncu_${name}_raw.csv
  rm -f gpurun_out/prof_$name.ncu-rep
done
grep -i "tensor pipe" gpurun_out/r02_ncu_gemm_fwd.md | head -8
du -sh gpurun_out
