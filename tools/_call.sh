cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode.py -x -q -m gpu > gpurun_out/c19_tests.log 2>&1; echo tests rc=$?; tail -12 gpurun_out/c19_tests.log | cut -c1-300
bash tools/ncu_capture.sh fp32 > gpurun_out/c18_capture_fp32.log 2>&1; tail -12 gpurun_out/c18_capture_fp32.log
