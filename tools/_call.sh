# Builder-side validation of a build on a B200 box:  gpurun --timeout 900 -- 'bash tools/_call.sh'
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1; echo tests rc=$? t=$(( $(date +%s) - T0 )); tail -2 gpurun_out/tests.log | cut -c1-200
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 420 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo bench rc=$? t=$(( $(date +%s) - T0 )); cut -c1-200 gpurun_out/bench.json
