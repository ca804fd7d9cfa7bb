cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/c17_tests.log 2>&1; echo tests rc=$?; tail -6 gpurun_out/c17_tests.log | cut -c1-300
timeout 400 python bench.py --images 200 --steps 2 --no-cpu-baseline --no-sub --profile-out gpurun_out/c17_layers_fp32.csv > gpurun_out/c17_bench_fp32.log 2>&1; echo bench rc=$?; tail -1 gpurun_out/c17_bench_fp32.log | cut -c1-120; head -2 gpurun_out/c17_layers_fp32.csv | tail -1
python -c "import __graft_entry__ as g; g.smoke()"
