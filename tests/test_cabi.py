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


@pytest.mark.parametrize("hw", [(1, 1), (1, 300), (5, 3), (100, 400), (112, 113), (113, 112), (150, 260), (223, 224),
                                (224, 225), (300, 340), (460, 230)])
def test_upload_window_holds_every_pixel_of_the_patch(hw):
    """mc_upload_window (the data-movement rule of mc_extract_images_host for sparsely annotated images): cropping the
    window around the re-expressed point gives the bytes of the whole-image crop -- checked with the oracle's crop_patches
    for the centres where reflection happens (all four borders and corners, +-2 around the 112 / 113 thresholds) and a
    grid of interior ones, on images smaller than, equal to and larger than a patch."""
    from oracle import crop as ocrop

    H, W = hw
    lib = _lib.load()
    rng = np.random.default_rng(H * 1000 + W)
    im = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)

    def axis(n):
        c = {0, 1, 2, n - 1, n - 2, n - 3, n // 2}
        for t in (110, 111, 112, 113, 114):
            c.update({t, n - 1 - t})
        return sorted(x for x in c if 0 <= x < n)

    pts = [(r, c) for r in axis(H) for c in axis(W)]
    want = ocrop.crop_patches(im, pts)
    out = (C.c_int32 * 4)()
    ptr = [C.cast(C.byref(out, 4 * i), C.c_void_p) for i in range(4)]
    for k, (r, c) in enumerate(pts):
        _lib.check(lib.mc_upload_window(H, W, r, c, *ptr))
        r0, c0, h, w = out[0], out[1], out[2], out[3]
        assert 0 <= r0 <= r < r0 + h <= H and 0 <= c0 <= c < c0 + w <= W
        assert h <= 225 and w <= 225
        got = ocrop.crop_patches(np.ascontiguousarray(im[r0:r0 + h, c0:c0 + w]), [(r - r0, c - c0)])[0]
        assert np.array_equal(got, want[k]), (hw, (r, c), (r0, c0, h, w))
    with pytest.raises(_lib.RowColumnInvalidError):
        _lib.check(lib.mc_upload_window(H, W, H, 0, *ptr))


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
