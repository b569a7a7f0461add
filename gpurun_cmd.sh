cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gn_bwd_resident_kernel -c 2 -f -o gpurun_out/ncu57_gnbwd python tools/profile_step.py --batch 512 > gpurun_out/ncu57_gnbwd.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/ncu57*
