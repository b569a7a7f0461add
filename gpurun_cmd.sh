cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "groupnorm" > gpurun_out/t46_gn.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/t46_gn.log
for m in 1 2 4 1 2; do
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --sample-steps 8 --micro $m > gpurun_out/bench46_m${m}.json 2> gpurun_out/bench46_m${m}.err; echo "bench micro=$m rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench46_m${m}.json'));print(d['ms_per_step'], d['value'], d['e2e']['value'], d['sampler'].get('value'), d['roofline']['achieved'], d['roofline']['gemm_ms_per_step'])"
done
