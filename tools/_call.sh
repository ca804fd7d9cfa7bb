cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
CMD0="python tools/fused_check.py --mask 2 --reps 1"
CMD3="python tools/fused_check.py --mask 10 --reps 1"
$CMD0 > gpurun_out/c2_plain0.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mbconv_fused -s 1 -c 1 -f -o gpurun_out/c2_fused0 $CMD0 > gpurun_out/c2_ncu0.log 2>&1
echo rc0=$?
$CMD3 > gpurun_out/c2_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mbconv_fused -s 1 -c 1 -f -o gpurun_out/c2_fused3 $CMD3 > gpurun_out/c2_ncu3.log 2>&1
echo rc3=$?
ls -la gpurun_out/*.ncu-rep
