"""Oracle self-checks for crop / normalize (CPU)."""
import numpy as np
import pytest
import torch

from oracle import crop


@pytest.mark.parametrize("hw", [(300, 400), (224, 224), (100, 37), (5, 3), (1, 7), (2, 2), (113, 500)])
def test_gather_equals_padded_slice(hw):
    H, W = hw
    rng = np.random.default_rng(H * 1000 + W)
    im = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    rc = [(0, 0), (H - 1, W - 1), (0, W - 1), (H - 1, 0), (H // 2, W // 2), (min(H - 1, 111), min(W - 1, 112))]
    a = crop.crop_patches_padded(im, rc)
    b = crop.crop_patches(im, rc)
    for x, y in zip(a, b):
        assert x.shape == (224, 224, 3)
        assert np.array_equal(x, y)


def test_centre_pixel_is_the_point():
    rng = np.random.default_rng(1)
    im = rng.integers(0, 256, (500, 700, 3), dtype=np.uint8)
    rc = [(0, 0), (499, 699), (250, 350), (111, 112), (113, 3)]
    p = crop.crop_patches(im, rc)
    for k, (r, c) in enumerate(rc):
        assert np.array_equal(p[k, 112, 112], im[r, c])


def test_reflect_index_single_reflection_formula():
    n = 300
    t = np.arange(-224, n + 224)
    want = np.where(t < 0, -t, np.where(t >= n, 2 * (n - 1) - t, t))
    assert np.array_equal(crop.reflect_index(t, n), want)


def test_normalize_matches_torchvision_formula():
    rng = np.random.default_rng(2)
    p = rng.integers(0, 256, (3, 224, 224, 3), dtype=np.uint8)
    got = crop.normalize_patches(p)
    t = torch.from_numpy(p).permute(0, 3, 1, 2).contiguous().to(torch.float32).div(255)
    mean = torch.tensor(crop.IMAGENET_MEAN, dtype=torch.float32)[None, :, None, None]
    std = torch.tensor(crop.IMAGENET_STD, dtype=torch.float32)[None, :, None, None]
    want = t.sub_(mean).div_(std).numpy()
    assert got.dtype == np.float32 and got.shape == (3, 3, 224, 224)
    assert np.array_equal(got, want)
    try:
        from PIL import Image
        from torchvision import transforms

        tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize(crop.IMAGENET_MEAN, crop.IMAGENET_STD)])
        tv = torch.stack([tf(Image.fromarray(q)) for q in p]).numpy()
        assert np.array_equal(got, tv)
    except ImportError:
        pass


def test_check_extract_inputs():
    crop.check_extract_inputs(100, 200, [(0, 0), (99, 199)])
    with pytest.raises(crop.RowColumnInvalidError):
        crop.check_extract_inputs(100, 200, [(100, 0)])
    with pytest.raises(crop.RowColumnInvalidError):
        crop.check_extract_inputs(100, 200, [(0, -1)])
    with pytest.raises(crop.DataLimitError):
        crop.check_extract_inputs(10001, 10000, [(0, 0)])
    with pytest.raises(crop.DataLimitError):
        crop.check_extract_inputs(100, 200, [(0, 0)] * 1001)


def test_bilinear_resize_oracle_is_torch_bit_for_bit():
    """The patch-size != 224 path (``model.json`` ``config.patch_size``, reference ``inference/export.py:77``): the NumPy
    restatement of the resize equals ``torch.nn.functional.interpolate(mode="bilinear", align_corners=False)`` on the float
    patch -- every float, hence every rounded byte -- for down- and up-scaling, dyadic and non-dyadic ratios."""
    import torch
    import torch.nn.functional as F

    from oracle.crop import _fma32, bilinear_taps, crop_patches, crop_resize_patches, resize_patches_bilinear

    rng = np.random.default_rng(0)
    for P in (448, 300, 226, 112, 100, 64):
        p = rng.integers(0, 256, (2, P, P, 3), dtype=np.uint8)
        ref = F.interpolate(torch.from_numpy(p).permute(0, 3, 1, 2).float(), size=(224, 224), mode="bilinear",
                            align_corners=False).permute(0, 2, 3, 1).numpy()
        got = resize_patches_bilinear(p)
        assert got.dtype == np.uint8 and got.shape == (2, 224, 224, 3)
        assert np.array_equal(got, np.clip(np.rint(ref), 0, 255).astype(np.uint8)), P
        i0, i1, l0, l1 = bilinear_taps(P, 224)
        assert i0.min() >= 0 and i1.max() == P - 1 and np.all(l0 + l1 == 1)
    im = rng.integers(0, 256, (90, 130, 3), dtype=np.uint8)
    rcs = [(0, 0), (89, 129), (40, 60), (3, 127)]
    assert np.array_equal(crop_resize_patches(im, rcs, 224), crop_patches(im, rcs))
    out = crop_resize_patches(im, rcs, 112)
    assert out.shape == (4, 224, 224, 3)
    # 2x up-sampling of the 112 window: output pixel (1, 1) sits at source (0.25, 0.25) of the window
    w = crop_patches(im, rcs, 112)[2].astype(np.float32)
    want = 0.75 * (0.75 * w[0, 0] + 0.25 * w[0, 1]) + 0.25 * (0.75 * w[1, 0] + 0.25 * w[1, 1])
    assert np.all(np.abs(out[2, 1, 1].astype(np.float32) - want) <= 0.5 + 1e-4)
    with pytest.raises(ValueError):
        crop_resize_patches(im, rcs, 225)
    assert _fma32(np.float32(3), np.float32(0.1), np.float32(1)) == np.float32(np.float64(np.float32(0.1)) * 3 + 1)
