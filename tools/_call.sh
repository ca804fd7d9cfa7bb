cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python tools/parity_stats.py > gpurun_out/c9_parity.log 2>&1; echo rc=$?; tail -5 gpurun_out/c9_parity.log
