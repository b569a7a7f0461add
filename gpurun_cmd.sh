cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 240 python tools/likelihood_bench.py 256 1e-3 > gpurun_out/lik58.txt 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/lik58.txt | cut -c1-900
