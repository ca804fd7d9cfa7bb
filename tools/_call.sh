cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_extract.py tests/test_gpu_drivers.py -x -q -m gpu -k "not 10k" > gpurun_out/c32_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/c32_tests.log | cut -c1-300
timeout 400 python bench.py --images 200 --steps 2 --no-cpu-baseline --no-sub --profile-out gpurun_out/c32_layers_fp32.csv > gpurun_out/c32_bench_fp32.log 2>&1; echo bench rc=$?; tail -1 gpurun_out/c32_bench_fp32.log | cut -c1-120; head -2 gpurun_out/c32_layers_fp32.csv | tail -1
python - <<'PY'
import time, numpy as np, torch, sys
sys.path.insert(0,'.')
from mermaid_classifier_b200 import synth
from mermaid_classifier_b200.extractor import EfficientNetExtractor
ext = EfficientNetExtractor(state_dict=synth.synth_backbone_state_dict(), mode="fp32", max_batch=1000)
rng = np.random.default_rng(0)
patches = rng.integers(0, 255, (1000, 224, 224, 3), dtype=np.uint8)
dev = torch.from_numpy(patches).cuda(); out = torch.empty((1000,1280), device='cuda')
from mermaid_classifier_b200 import _lib
h = ext._ensure_handle(); lib = _lib.load()
for _ in range(2): _lib.check(lib.mc_extract_patches(h, dev.data_ptr(), 1000, out.data_ptr(), _lib.stream_ptr()))
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(5): _lib.check(lib.mc_extract_patches(h, dev.data_ptr(), 1000, out.data_ptr(), _lib.stream_ptr()))
torch.cuda.synchronize(); print("pre-cropped patches/s:", 5000/(time.perf_counter()-t0))
PY
