cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "groupnorm" > gpurun_out/t52_gn.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/t52_gn.log
timeout 300 python tools/gn_bench.py > gpurun_out/gn_bench52.txt 2>&1; echo "gnbench rc=$?"; grep "resident x[0-9]*:\|^[0-9]" gpurun_out/gn_bench52.txt | cut -c1-200
for i in 1 2; do for f in 0 1; do
ST_GN_BWD_RESIDENT=$f timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --sample-steps 4 > gpurun_out/bench52_r${f}_$i.json 2> gpurun_out/bench52_r${f}_$i.err; echo "bench resident=$f rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench52_r${f}_$i.json'));print(d['ms_per_step'], d['value'])"
done; done
