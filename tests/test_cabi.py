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


@pytest.mark.parametrize("hw,n,spread", [((700, 900), 40, None), ((700, 900), 60, 150), ((260, 1200), 25, None),
                                         ((100, 500), 12, None), ((1500, 2000), 80, 400)])
def test_upload_plan_merges_windows_and_keeps_every_patch(hw, n, spread):
    """mc_plan_uploads (what mc_extract_images_host uploads for a sparsely annotated image): the points' windows merged into
    bounding boxes.  Every point's own window lies inside its planned window, the patch cropped from the planned window equals
    the oracle's patch of the whole image byte for byte, clustered points share windows, and a merge never costs more than the
    stated slack (128 KiB) beyond the point's own window."""
    from oracle import crop as ocrop

    H, W = hw
    lib = _lib.load()
    rng = np.random.default_rng(H + 7 * W + n)
    im = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    if spread is None:
        rc = np.stack([rng.integers(0, H, n), rng.integers(0, W, n)], 1)
    else:   # clustered annotations: overlapping windows
        centre = np.array([H // 3, W // 3])
        rc = np.clip(centre + rng.integers(-spread, spread + 1, size=(n, 2)), 0, [H - 1, W - 1])
    rc[:4] = [(0, 0), (0, W - 1), (H - 1, 0), (H - 1, W - 1)]
    rc = np.ascontiguousarray(rc, dtype=np.int32)
    wins = np.zeros((n, 4), np.int32)
    idx = np.zeros(n, np.int32)
    nw = C.c_int64()
    _lib.check(lib.mc_plan_uploads(H, W, rc.ctypes.data, n, wins.ctypes.data, idx.ctypes.data, C.byref(nw)))
    nw = nw.value
    assert 1 <= nw <= n and idx.min() >= 0 and idx.max() == nw - 1
    if spread is not None:
        assert nw < n // 2          # clustered points share windows
    want = ocrop.crop_patches(im, [tuple(p) for p in rc])
    own = (C.c_int32 * 4)()
    ptr = [C.cast(C.byref(own, 4 * i), C.c_void_p) for i in range(4)]
    own_bytes = 0
    for k, (r, c) in enumerate(rc):
        r0, c0, h, w = (int(v) for v in wins[idx[k]])
        _lib.check(lib.mc_upload_window(H, W, int(r), int(c), *ptr))
        assert r0 <= own[0] and c0 <= own[1] and r0 + h >= own[0] + own[2] and c0 + w >= own[1] + own[3]
        assert 0 <= r0 and r0 + h <= H and 0 <= c0 and c0 + w <= W
        own_bytes += own[2] * own[3] * 3
        got = ocrop.crop_patches(np.ascontiguousarray(im[r0:r0 + h, c0:c0 + w]), [(int(r) - r0, int(c) - c0)])[0]
        assert np.array_equal(got, want[k]), (hw, (r, c), (r0, c0, h, w))
    planned = int((wins[:nw, 2].astype(np.int64) * wins[:nw, 3] * 3).sum())
    assert planned <= own_bytes + (n - nw) * 128 * 1024


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
