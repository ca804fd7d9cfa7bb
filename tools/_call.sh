cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_callers.py tests/test_gpu_mlp.py tests/test_gpu_trainer.py -x -q -m gpu > gpurun_out/c22_tests.log 2>&1; echo tests rc=$?; tail -8 gpurun_out/c22_tests.log | cut -c1-300
