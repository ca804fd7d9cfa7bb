#!/bin/bash
# Round-1 ncu evidence (run under gpurun): launch list of one 500-patch sub-batch + full-set captures of one
# kernel per family.  Outputs under gpurun_out/; tools/ncu_table.py / ncu_src.py summarise them into profiles/.
MODE=${1:-fp32}
CMD="python bench.py --images 5 --batch 500 --steps 1 --no-cpu-baseline --mode $MODE"
timeout 120 $CMD > gpurun_out/plain_$MODE.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$MODE.log; exit 1; }
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$MODE.csv $CMD > gpurun_out/ncu_l.log 2>&1
cap() {  # name regex skip
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o gpurun_out/prof_${1}_$MODE -f $CMD > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
}
cap stem stem_kernel 0
cap dw_b1 dw_tma_kernel 1
cap dw_b4 dw_reg_kernel 2
cap exp_b1 "pw_tc_kernel.*Lb0" 0
cap proj_b2 "pw_tc_kernel.*Lb1" 2
cap head_conv "pw_tc_kernel.*Lb0" 15
cap head_rows head_rows_kernel 0
ls -la gpurun_out/*.ncu-rep
