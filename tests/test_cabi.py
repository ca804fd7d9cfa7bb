"""CPU checks of the boundary: the shared library builds, loads and exports every symbol the
header declares; host-only entry points behave; compute entry points fail loudly without a GPU."""
import ctypes as C

import numpy as np
import pytest
import torch

from mermaid_classifier_b200 import _lib, weights


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _lib.header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"libmermaid_b200.so lacks {name}"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == declared
    assert lib.mc_abi_version() == 1


def test_param_count_matches_packer(backbone_sd):
    lib = _lib.load()
    blob = weights.pack_backbone(backbone_sd)
    assert blob.dtype == np.float32
    assert blob.size == lib.mc_backbone_param_count()


def test_check_extract_inputs_host_only():
    lib = _lib.load()
    rc = np.array([[0, 0], [99, 199]], dtype=np.int32)
    assert lib.mc_check_extract_inputs(100, 200, rc.ctypes.data, 2, 0, 0) == 0
    bad = np.array([[100, 0]], dtype=np.int32)
    with pytest.raises(_lib.RowColumnInvalidError):
        _lib.check(lib.mc_check_extract_inputs(100, 200, bad.ctypes.data, 1, 0, 0))
    with pytest.raises(_lib.DataLimitError):
        _lib.check(lib.mc_check_extract_inputs(10001, 10000, rc.ctypes.data, 2, 0, 0))
    with pytest.raises(_lib.DataLimitError):
        _lib.check(lib.mc_check_extract_inputs(100, 200, rc.ctypes.data, 1001, 0, 0))


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback(backbone_sd):
    from mermaid_classifier_b200.extractor import EfficientNetExtractor

    e = EfficientNetExtractor(state_dict=backbone_sd)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        e.extract_array(np.zeros((300, 300, 3), dtype=np.uint8), [(1, 1)])
    lib = _lib.load()
    blob = weights.pack_backbone(backbone_sd)
    h = C.c_void_p()
    status = lib.mc_extractor_create(blob.ctypes.data, blob.size, 0, 0, 8, C.byref(h))
    assert status == _lib.MC_ERR_CUDA and b"no CPU fallback" in lib.mc_last_error()


def test_serving_imports_stay_light():
    """Mirror of the reference's inference/training split (tests/pyspacer/test_inference_decoupling.py): importing the
    serving-side modules must not drag in the training-only stack (scikit-learn, scipy, pyspacer) -- checked in a fresh
    interpreter so modules imported by this session cannot mask a regression."""
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parents[1]
    for target in ("mermaid_classifier_b200.inference", "mermaid_classifier_b200.extractor", "mermaid_classifier_b200.export"):
        script = (f"import sys\nsys.path.insert(0, {str(root)!r})\nimport {target}\n"
                  "heavy = [m for m in ('sklearn', 'scipy', 'spacer', 'pandas') if m in sys.modules]\n"
                  "assert not heavy, heavy\nprint('ok')\n")
        res = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True)
        assert res.returncode == 0 and "ok" in res.stdout, (target, res.stderr[-400:])
