"""Device-side image decode feeding the crop kernel.

The reference loads every image with ``spacer.storage.load_image`` -- PIL decodes the JPEG on the CPU and converts it to
RGB (call site ``/root/reference/mermaid_classifier/pyspacer/annotation.py:235``; inside ``spacer.tasks.extract_features``,
``/root/reference/scripts/build_feature_bucket.py:775``) -- and the decoded 36 MB image is what travels on.  Here the
COMPRESSED bytes go to the library (``mc_jpeg_decode``: Huffman decoding on the calling thread, IDCT / upsampling / colour
conversion on the GPU through nvJPEG) and come out as an RGB8 device image that ``EfficientNetExtractor.extract_device``
reads directly; the decoded image never crosses PCIe.

Two device paths.  ``exact=True`` (default) is ``mc_jpeg_decode_exact``: the library's own baseline + progressive decoder -- entropy
decoding on the calling thread into a sparse coefficient stream, then libjpeg-turbo's integer IDCT, fancy chroma upsampling
and YCbCr -> RGB arithmetic restated as kernels -- whose bytes EQUAL PIL's (``tests/test_gpu_decode.py``; the CPU restatement
``oracle/jpeg.py`` is pinned to PIL byte for byte by ``tests/test_oracle_jpeg.py``).  Streams it does not cover (CMYK,
arithmetic coding, unusual sampling factors) fall through to nvJPEG (``mc_jpeg_decode``), whose inverse DCT and chroma interpolation are
not libjpeg-turbo's bit for bit: against PIL its bytes differ by a few grey levels (the bound is stated and checked in the same
test file).  Files that are not JPEG streams (the PNG stand-ins of the tests) fall back to PIL + one host-to-device copy.
"""

from __future__ import annotations

import ctypes as C
import io
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Any, Sequence

import numpy as np

from . import _lib


def is_jpeg(data: bytes) -> bool:
    return len(data) > 3 and data[0] == 0xFF and data[1] == 0xD8


def jpeg_coefficients(data: bytes):
    """Quantised DCT coefficients of a baseline JPEG stream as the library's host-side entropy decoder produces them
    (``mc_jpeg_coefficients_host``; no GPU involved): ``(info, [per-component (by, bx, 64) int16 arrays])`` with
    ``info = {"height", "width", "components", "restart_interval"}``."""
    data = bytes(data)
    info = (C.c_int32 * 10)()
    lib = _lib.load()
    _lib.check(lib.mc_jpeg_coefficients_host(data, len(data), None, 0, info))
    nc = info[2]
    grids = [(info[4 + 2 * c], info[3 + 2 * c]) for c in range(nc)]
    total = sum(by * bx for by, bx in grids)
    flat = np.zeros((total, 64), dtype=np.int16)
    _lib.check(lib.mc_jpeg_coefficients_host(data, len(data), flat.ctypes.data, total, info))
    out, pos = [], 0
    for by, bx in grids:
        out.append(flat[pos: pos + by * bx].reshape(by, bx, 64))
        pos += by * bx
    return {"height": info[0], "width": info[1], "components": nc, "restart_interval": info[9]}, out


class JpegDecoder:
    """One ``mc_jpeg`` handle (not thread-safe: one per worker thread)."""

    def __init__(self, device: int | None = None, exact: bool = True):
        torch = _lib.require_cuda()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.exact = bool(exact)
        self.last_path = None   # "exact" / "nvjpeg": which decoder produced the last image
        h = C.c_void_p()
        _lib.check(_lib.load().mc_jpeg_create(self.device, C.byref(h)))
        self._h = h

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            _lib.load().mc_jpeg_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def info(self, data: bytes) -> tuple[int, int, int]:
        """``(height, width, components)`` of a JPEG stream."""
        h, w, nc = C.c_int32(), C.c_int32(), C.c_int32()
        data = bytes(data)
        _lib.check(_lib.load().mc_jpeg_info(self._h, data, len(data), C.byref(h), C.byref(w), C.byref(nc)))
        return h.value, w.value, nc.value

    def decode(self, data: bytes, stream: Any = None):
        """JPEG bytes -> CUDA ``(H, W, 3) uint8`` tensor (RGB).  Asynchronous on ``stream`` (default: current)."""
        torch = _lib.require_cuda()
        data = bytes(data)
        H, W, _ = self.info(data)
        with torch.cuda.device(self.device):
            out = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
            if self.exact:
                status = _lib.load().mc_jpeg_decode_exact(self._h, data, len(data), out.data_ptr(), W * 3, H, W, _lib.stream_ptr(stream))
                if status == _lib.MC_OK:
                    self.last_path = "exact"
                    return out
                if status != _lib.MC_ERR_UNSUPPORTED:
                    _lib.check(status)
            _lib.check(_lib.load().mc_jpeg_decode(self._h, data, len(data), out.data_ptr(), W * 3, H, W, _lib.stream_ptr(stream)))
            self.last_path = "nvjpeg"
        return out


def load_image_device(data: bytes, decoder: JpegDecoder, stream: Any = None):
    """``spacer.storage.load_image`` semantics on the device: RGB8 HWC CUDA tensor from an encoded image.  JPEG streams are
    decoded by the library; anything else goes through PIL (``convert("RGB")``) and one host-to-device copy."""
    torch = _lib.require_cuda()
    if is_jpeg(data):
        return decoder.decode(data, stream)
    from PIL import Image

    arr = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
    with torch.cuda.device(decoder.device):
        return torch.from_numpy(np.ascontiguousarray(arr)).cuda(non_blocking=False)


class DecodePool:
    """``n_threads`` workers, each with its own decoder handle and CUDA stream: the host part of nvJPEG (Huffman decoding)
    runs in parallel, the GPU parts overlap on the streams.  ``decode_many`` returns device images in input order, ready
    on the caller's current stream."""

    def __init__(self, n_threads: int = 8, device: int | None = None, exact: bool = True):
        torch = _lib.require_cuda()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.exact = bool(exact)
        self.n_threads = max(1, int(n_threads))
        self._pool = ThreadPoolExecutor(max_workers=self.n_threads)
        self._local = threading.local()
        self._decoders: list[JpegDecoder] = []
        self._lock = threading.Lock()

    def _worker_state(self):
        torch = _lib.require_cuda()
        st = getattr(self._local, "state", None)
        if st is None:
            dec = JpegDecoder(self.device, exact=self.exact)
            with self._lock:
                self._decoders.append(dec)
            st = self._local.state = (dec, torch.cuda.Stream(device=self.device))
        return st

    def _one(self, data: bytes):
        torch = _lib.require_cuda()
        dec, stream = self._worker_state()
        try:
            with torch.cuda.device(self.device), torch.cuda.stream(stream):
                img = load_image_device(data, dec, stream)
                ev = torch.cuda.Event()
                ev.record(stream)
            return img, ev, None
        except KeyboardInterrupt:
            raise
        except Exception as exc:
            return None, None, exc

    def decode_many(self, blobs: Sequence[bytes]) -> list[tuple[Any, Exception | None]]:
        """``[(image_or_None, exception_or_None), ...]`` in input order."""
        torch = _lib.require_cuda()
        out = []
        cur = torch.cuda.current_stream(self.device)
        for img, ev, exc in self._pool.map(self._one, blobs):
            if ev is not None:
                cur.wait_event(ev)
                img.record_stream(cur)
            out.append((img, exc))
        return out

    def close(self) -> None:
        self._pool.shutdown(wait=True)
        for d in self._decoders:
            d.close()
        self._decoders = []
