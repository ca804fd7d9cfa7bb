cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/v2_tests.log 2>&1; echo tests rc=$? t=$(( $(date +%s) - T0 )); tail -2 gpurun_out/v2_tests.log | cut -c1-200
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 420 python bench.py > gpurun_out/v2_bench.json 2> gpurun_out/v2_bench.err; echo bench rc=$? t=$(( $(date +%s) - T0 )); cut -c1-300 gpurun_out/v2_bench.json
timeout 200 python bench.py --mode bf16 --images 300 --no-cpu-baseline > gpurun_out/v2_bench_bf16.json 2>/dev/null; echo bf16 rc=$? t=$(( $(date +%s) - T0 )); cut -c1-200 gpurun_out/v2_bench_bf16.json
timeout 100 python bench.py --mode fp32 --images 100 --no-cpu-baseline --no-sub --profile-out gpurun_out/v2_layers_fp32.csv > /dev/null 2>&1
timeout 100 python bench.py --mode bf16 --images 100 --no-cpu-baseline --no-sub --profile-out gpurun_out/v2_layers_bf16.csv > /dev/null 2>&1
echo t=$(( $(date +%s) - T0 ))
