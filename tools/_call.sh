cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for m in 2 4 8; do timeout 120 python tools/fused_check.py --mask $m --images 10 > gpurun_out/c7_fused_$m.log 2>&1; echo fused $m rc=$?; tail -1 gpurun_out/c7_fused_$m.log | cut -c1-600; done
for m in 2 4 8; do timeout 120 python tools/fused_check.py --mode bf16 --mask $m --images 10 > gpurun_out/c7_fusedb_$m.log 2>&1; echo fused bf16 $m rc=$?; tail -1 gpurun_out/c7_fusedb_$m.log | cut -c1-600; done
