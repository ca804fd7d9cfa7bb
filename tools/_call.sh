cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export MC_NVCC_DEFS="-DMC_TC_TIMING=1"
for spec in bf16:1 fp32:1 bf16:3; do
  m=${spec%%:*}; l=${spec##*:}
  MC_TC_DBG=$l timeout 100 python bench.py --mode $m --images 20 --steps 1 --warmup 1 --no-cpu-baseline --no-sub 2>&1 | grep "MC_TC_DBG" | head -1 | sed "s/^/$m /"
done
