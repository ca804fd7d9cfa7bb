cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_extract.py tests/test_gpu_drivers.py -x -q -m gpu -k "extract_many or window or bucket" 2>&1 | tail -2
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('$1', 'value', round(d['value']), 'e2e', round(e['value']), 'ms/step', round(e['ms_per_step'],1), 'chk', e['labels_checksum'])"; }
MC_PIPE_DEBUG=1 timeout 200 python bench.py --mode bf16 --points 50 --images 300 --no-cpu-baseline --no-sub 2> gpurun_out/pipe_dbg.txt | show "bf16 C3 driver-api"
grep "mc pipe" gpurun_out/pipe_dbg.txt | tail -1
MC_H2D_RUNTIME_API=1 MC_PIPE_DEBUG=1 timeout 200 python bench.py --mode bf16 --points 50 --images 300 --no-cpu-baseline --no-sub 2> gpurun_out/pipe_dbg.txt | show "bf16 C3 runtime-api"
grep "mc pipe" gpurun_out/pipe_dbg.txt | tail -1
