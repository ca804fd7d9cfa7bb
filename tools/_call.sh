cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_extract.py tests/test_gpu_head.py -x -q -m gpu -k "not 10k" > gpurun_out/c35_tests.log 2>&1; echo tests rc=$?; tail -4 gpurun_out/c35_tests.log | cut -c1-300
timeout 400 python bench.py --images 200 --steps 2 --no-cpu-baseline --no-sub --profile-out gpurun_out/c35_layers_fp32.csv > gpurun_out/c35_bench_fp32.log 2>&1; echo bench rc=$?; tail -1 gpurun_out/c35_bench_fp32.log | cut -c1-120
timeout 400 python bench.py --images 200 --steps 2 --no-cpu-baseline --no-sub --mode bf16 --profile-out gpurun_out/c35_layers_bf16.csv > gpurun_out/c35_bench_bf16.log 2>&1; echo bench rc=$?; tail -1 gpurun_out/c35_bench_bf16.log | cut -c1-120
