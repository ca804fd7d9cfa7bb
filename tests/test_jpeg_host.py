"""The host half of ``mc_jpeg_decode_exact`` -- marker parsing and Huffman decoding in the C library
(``mc_jpeg_coefficients_host``, no GPU) -- against the oracle's entropy decoder, and, through the oracle's IDCT / upsampling /
colour stages, against PIL: the CPU suite thereby pins everything of the exact decoder that does not run on the device."""
import io

import numpy as np
import pytest
from PIL import Image

from mermaid_classifier_b200.decode import jpeg_coefficients
from oracle import jpeg as oj


def _photo(rng, h, w, noise=20.0):
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([128 + 90 * np.sin(xx / 9.0 + yy / 17.0), 128 + 80 * np.cos(xx / 13.0 - yy / 7.0),
                     128 + 70 * np.sin((xx + yy) / 11.0)], -1)
    return np.clip(base + rng.normal(0, noise, (h, w, 3)), 0, 255).astype(np.uint8)


def _jpeg(arr, **kw):
    buf = io.BytesIO()
    Image.fromarray(arr).save(buf, format="JPEG", **kw)
    return buf.getvalue()


def test_host_entropy_decoder_matches_oracle_and_pil():
    rng = np.random.default_rng(5)
    cases = [(_photo(rng, h, w), dict(quality=q, subsampling=ss))
             for (h, w) in [(48, 64), (45, 67), (8, 8), (1, 1), (100, 3), (31, 49)] for ss in (0, 1, 2) for q in (95, 40, 5)]
    sat = np.zeros((40, 56, 3), np.uint8)
    sat[::2] = 255
    cases += [(sat, dict(quality=85, subsampling=2, optimize=True)), (sat, dict(quality=85, subsampling=0, restart_marker_blocks=3)),
              (_photo(rng, 60, 90), dict(quality=60, subsampling=1, restart_marker_rows=1)), (_photo(rng, 33, 47)[:, :, 0], dict(quality=75))]
    for arr, kw in cases:
        data = _jpeg(arr, **kw)
        info, coefs = jpeg_coefficients(data)
        hdr = oj.parse(data)
        want = oj.decode_coefficients(data, hdr)
        assert (info["height"], info["width"], info["components"]) == (arr.shape[0], arr.shape[1], 1 if arr.ndim == 2 else 3)
        assert info["restart_interval"] == hdr["ri"]
        for a, b in zip(coefs, want):
            assert a.shape == b.shape and np.array_equal(a, b), kw
        rgb = oj.rgb_from_planes(oj.planes_from_coefficients([c.astype(np.int32) for c in coefs], hdr), hdr)
        assert np.array_equal(rgb, np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))), kw


def test_host_progressive_decoder_matches_oracle_and_pil():
    """Progressive streams (jdphuff.c: DC / AC first and refinement scans, EOB runs, per-scan tables, restarts)."""
    rng = np.random.default_rng(8)
    cases = [(_photo(rng, h, w), dict(quality=q, subsampling=ss, progressive=True, **extra))
             for (h, w) in [(48, 64), (45, 67), (8, 8), (1, 1), (100, 3), (5, 4), (31, 49)] for ss in (0, 1, 2)
             for q, extra in ((95, {}), (40, dict(optimize=True)), (8, dict(restart_marker_blocks=3)))]
    cases.append((_photo(rng, 33, 47)[:, :, 0], dict(quality=75, progressive=True)))
    for arr, kw in cases:
        data = _jpeg(arr, **kw)
        info, coefs = jpeg_coefficients(data)
        hdr = oj.parse(data)
        assert hdr["frame"]["progressive"] and (info["height"], info["width"]) == arr.shape[:2]
        want = oj.decode_coefficients_progressive(data, hdr)
        for a, b in zip(coefs, want):
            assert a.shape == b.shape and np.array_equal(a, b), kw
        rgb = oj.rgb_from_planes(oj.planes_from_coefficients([c.astype(np.int32) for c in coefs], hdr), hdr)
        assert np.array_equal(rgb, np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))), kw


def test_host_progressive_decoder_survives_damaged_streams():
    rng = np.random.default_rng(9)
    good = _jpeg(_photo(rng, 96, 128), quality=85, subsampling=2, progressive=True)
    for cut in range(300, len(good), 211):
        try:
            jpeg_coefficients(good[:cut])
        except (ValueError, RuntimeError):
            pass
    for _ in range(60):
        bad = bytearray(good)
        for pos in rng.integers(20, len(good) - 2, size=10):
            bad[pos] ^= 1 << int(rng.integers(0, 8))
        try:
            jpeg_coefficients(bytes(bad))
        except (ValueError, RuntimeError):
            pass


def test_host_entropy_decoder_rejects_what_it_does_not_cover():
    rng = np.random.default_rng(6)
    im = _photo(rng, 40, 40)
    cmyk = io.BytesIO()
    Image.fromarray(np.dstack([im, im[:, :, 0]]), mode="CMYK").save(cmyk, format="JPEG", quality=80)
    with pytest.raises(RuntimeError):   # MC_ERR_UNSUPPORTED: four components
        jpeg_coefficients(cmyk.getvalue())
    with pytest.raises(ValueError):     # MC_ERR_BAD_ARG
        jpeg_coefficients(b"\x89PNG\r\n not a jpeg stream")
    good = _jpeg(im, quality=80)
    info, coefs = jpeg_coefficients(good[: len(good) // 2])   # a truncated scan decodes (zero-padded), as libjpeg does
    assert info["height"] == 40 and coefs[0].shape == (6, 6, 64)   # 4:2:0: three 16-row MCUs
