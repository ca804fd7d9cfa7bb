cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_extract.py tests/test_gpu_drivers.py tests/test_gpu_callers.py -x -q -m gpu -k "not 10k" 2>&1 | tail -2
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('$1', 'value', round(d['value']), 'e2e', round(e['value']), 'h2d_GB', round(e['h2d_bytes_per_step']/1e9,2), 'ms/step', round(e['ms_per_step'],1), 'chk', e['labels_checksum'], d['labels_checksum'])"; }
MC_PIPE_DEBUG=1 timeout 100 python bench.py --mode bf16 --points 50 --images 300 --no-cpu-baseline --no-sub 2> gpurun_out/pipe_dbg.txt | show "bf16 C3 merged"
grep "mc pipe" gpurun_out/pipe_dbg.txt | tail -1
MC_PIPE_DEBUG=1 timeout 100 python bench.py --images 150 --no-cpu-baseline --no-sub 2> gpurun_out/pipe_dbg2.txt | show "fp32 C2 merged"
grep "mc pipe" gpurun_out/pipe_dbg2.txt | tail -1
