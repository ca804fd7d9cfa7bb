"""The CPU restatement of libjpeg-turbo's baseline decode (``oracle/jpeg.py``) against PIL itself -- the decoder behind the
reference's ``spacer.storage.load_image`` (call site ``mermaid_classifier/pyspacer/annotation.py:235``): every byte equal on
4:4:4 / 4:2:2 / 4:2:0 / grayscale streams, odd and tiny sizes (the narrow-component replication rule of ``jdsample.c``),
low and high quality, optimised Huffman tables and restart intervals."""
import io

import numpy as np
import pytest
from PIL import Image

from oracle import jpeg as oj


def _photo(rng, h, w, noise=20.0):
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([128 + 90 * np.sin(xx / 9.0 + yy / 17.0), 128 + 80 * np.cos(xx / 13.0 - yy / 7.0),
                     128 + 70 * np.sin((xx + yy) / 11.0)], -1)
    return np.clip(base + rng.normal(0, noise, (h, w, 3)), 0, 255).astype(np.uint8)


def _pil(data):
    return np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))


def _jpeg(arr, **kw):
    buf = io.BytesIO()
    Image.fromarray(arr).save(buf, format="JPEG", **kw)
    return buf.getvalue()


@pytest.mark.parametrize("subsampling", [0, 1, 2])
def test_oracle_decode_equals_pil(subsampling):
    rng = np.random.default_rng(subsampling)
    for h, w in [(48, 64), (45, 67), (17, 33), (8, 8), (1, 1), (100, 3), (3, 100), (5, 4), (31, 49)]:
        for q in (95, 75, 20, 5):
            data = _jpeg(_photo(rng, h, w), quality=q, subsampling=subsampling)
            assert np.array_equal(oj.decode_rgb(data), _pil(data)), (h, w, q)


def test_oracle_decode_variants():
    rng = np.random.default_rng(7)
    im = _photo(rng, 40, 56)
    sat = np.zeros((40, 56, 3), np.uint8)
    sat[::2] = 255
    sat[:, ::3, 1] = 0
    for arr, kw in [(im[:, :, 0], dict(quality=80)), (sat, dict(quality=85, subsampling=2, optimize=True)),
                    (sat, dict(quality=85, subsampling=0, restart_marker_blocks=3)),
                    (im, dict(quality=60, subsampling=1, restart_marker_rows=1)), (sat, dict(quality=100, subsampling=0))]:
        data = _jpeg(arr, **kw)
        assert np.array_equal(oj.decode_rgb(data), _pil(data)), kw
    with pytest.raises(oj.UnsupportedJpeg):
        oj.decode_rgb(b"\x89PNG not a jpeg")
    # the pieces, on their own edge cases
    assert np.array_equal(oj.h2v1_fancy(np.array([[10, 20, 30]], np.uint8)), [[10, 13, 17, 23, 27, 30]])
    assert oj.range_limit_idct(np.array([-600, -129, -128, 0, 127, 128, 600])).tolist() == [255, 0, 0, 128, 255, 255, 0]   # the table wraps outside [-512, 511]


@pytest.mark.parametrize("subsampling", [0, 1, 2])
def test_oracle_progressive_decode_equals_pil(subsampling):
    """``jdphuff.c``: DC / AC first and refinement scans (spectral selection + successive approximation, EOB runs, restart
    intervals) accumulate to the same coefficients, hence the same bytes as PIL."""
    rng = np.random.default_rng(20 + subsampling)
    for h, w in [(48, 64), (45, 67), (17, 33), (8, 8), (1, 1), (100, 3), (31, 49)]:
        for q in (95, 60, 10):
            data = _jpeg(_photo(rng, h, w), quality=q, subsampling=subsampling, progressive=True)
            assert oj.parse(data)["frame"]["progressive"]
            assert np.array_equal(oj.decode_rgb(data), _pil(data)), (h, w, q)
    im = _photo(rng, 40, 56)
    for arr, kw in [(im[:, :, 0], dict(quality=80, progressive=True)),
                    (im, dict(quality=70, subsampling=subsampling, progressive=True, restart_marker_rows=1))]:
        data = _jpeg(arr, **kw)
        assert np.array_equal(oj.decode_rgb(data), _pil(data)), kw
