cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
PYTHONPATH=. timeout 200 python tools/sparse_check.py 2>&1 | tail -12
timeout 300 python -m pytest tests/test_gpu_extract.py tests/test_gpu_drivers.py tests/test_gpu_callers.py -x -q -m gpu -k "not 10k" 2>&1 | tail -3
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('$1', 'value', round(d['value']), 'e2e', round(e['value']), 'h2d_GB', round(e['h2d_bytes_per_step']/1e9,2), 'chk', e['labels_checksum'], d['labels_checksum'])"; }
for sp in 1 0; do
  MC_SPARSE_H2D=$sp timeout 200 python bench.py --images 100 --no-cpu-baseline --no-sub 2>/dev/null | show "fp32 C2 sparse=$sp"
done
