cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_head.py tests/test_gpu_callers.py tests/test_gpu_drivers.py tests/test_gpu_trainer.py -x -q -m gpu > gpurun_out/c36_tests.log 2>&1; echo tests rc=$?; tail -4 gpurun_out/c36_tests.log | cut -c1-300
python tools/bench_head.py --rows 4000000 --check 200000 2>&1 | tail -1 | cut -c1-400
