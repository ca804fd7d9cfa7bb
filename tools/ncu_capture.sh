#!/bin/bash
# ncu evidence (run under gpurun): launch list of one 500-patch sub-batch + `--set full` captures of one kernel
# per family.  Reports are summarised ON the GPU box (tools/ncu_table.py) and deleted: gpurun_out/ is capped at 64 MiB.
MODE=${1:-fp32}; shift
# round 2: b1.expand + b1.depthwise are one kernel (mbconv_fused), the stem runs on the tensor cores, the head conv pools in its epilogue
# fp32 defaults: b1-b3 run as mbconv_fused launches 0-2, dw_reg starts at b4, and the pw_tc launches of a sub-batch are
# b0.project(0) b1.project(1) b2.project(2) b3.project(3) b4.expand(4) b4.project(5) ... bK.expand(4 + 2 (K - 4)) bK.project(5 + 2 (K - 4)) ... head conv + pool(28)
CAPS=${@:-"stem:stem_tc_kernel:0 fused_b1:mbconv_fused_kernel:0 fused_b2:mbconv_fused_kernel:1 dw_b0:dw_tma_kernel:0 dw_b4:dw_reg_kernel:0 dw_b9:dw_reg_kernel:5 exp_b4:pw_tc_kernel:4 exp_b9:pw_tc_kernel:14 exp_b12:pw_tc_kernel:20 proj_b2:pw_tc_kernel:2 proj_b12:pw_tc_kernel:21 proj_b15:pw_tc_kernel:27 head_pool:pw_tc_kernel:28"}
CMD="python bench.py --images 5 --batch 500 --steps 1 --no-cpu-baseline --no-sub --mode $MODE"
timeout 120 $CMD > gpurun_out/plain_$MODE.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$MODE.log; exit 1; }
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$MODE.csv $CMD > gpurun_out/ncu_l.log 2>&1
OUT=gpurun_out/kernels_$MODE.csv; : > $OUT
for spec in $CAPS; do
  IFS=: read name regex skip <<< "$spec"
  timeout 200 ncu --set full --clock-control none -k regex:$regex -s $skip -c 1 -o /tmp/prof_$name -f $CMD > gpurun_out/ncu_$name.log 2>&1
  if [ -f /tmp/prof_$name.ncu-rep ]; then
    python tools/ncu_table.py /tmp/prof_$name.ncu-rep | awk -v n=$name 'NR==1 && !h {print "capture," $0} NR>1 {print n "," $0}' h=$( [ -s $OUT ] && echo 1 ) >> $OUT
    rm -f /tmp/prof_$name.ncu-rep
  else
    echo "capture $name failed"; tail -2 gpurun_out/ncu_$name.log
  fi
done
cat $OUT | cut -c1-160
