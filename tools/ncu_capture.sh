#!/bin/bash
# ncu evidence (run under gpurun): launch list of one 500-patch sub-batch + `--set full` captures of one kernel
# per family.  Reports are summarised ON the GPU box (tools/ncu_table.py) and deleted: gpurun_out/ is capped at 64 MiB.
MODE=${1:-fp32}; shift
# round 2: b1.expand + b1.depthwise are one kernel (mbconv_fused), the stem runs on the tensor cores, the head conv pools in its epilogue
CAPS=${@:-"stem:stem_tc_kernel:0 fused_b1:mbconv_fused_kernel:0 dw_b0:dw_tma_kernel:0 dw_b4:dw_reg_kernel:2 dw_b9:dw_reg_kernel:7 exp_b2:pw_tc_kernel:2 exp_b9:pw_tc_kernel:16 exp_b12:pw_tc_kernel:22 proj_b2:pw_tc_kernel:3 proj_b12:pw_tc_kernel:23 proj_b15:pw_tc_kernel:29 head_pool:pw_tc_kernel:30"}
CMD="python bench.py --images 5 --batch 500 --steps 1 --no-cpu-baseline --no-sub --mode $MODE"
timeout 120 $CMD > gpurun_out/plain_$MODE.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$MODE.log; exit 1; }
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$MODE.csv $CMD > gpurun_out/ncu_l.log 2>&1
# pw_tc launches of a sub-batch in order: b0.project(0) b1.project(1) b2.expand(2) b2.project(3) ... b15.expand(28) b15.project(29) head conv + pool(30)
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
