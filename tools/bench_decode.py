"""Device JPEG decode throughput (mc_jpeg_decode_exact and mc_jpeg_decode through DecodePool) on a 4000x3000 photograph-like
image, against PIL."""
import io, json, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from PIL import Image
from mermaid_classifier_b200 import synth
from mermaid_classifier_b200.decode import DecodePool

base = synth.synth_image(synth.DEFAULT_SEED, 1, 375, 500)
img = np.kron(base, np.ones((8, 8, 1), np.uint8)).astype(np.int16)
img = np.clip(img + np.random.default_rng(0).integers(-6, 7, img.shape), 0, 255).astype(np.uint8)
buf = io.BytesIO()
Image.fromarray(img).save(buf, format="JPEG", quality=90)
data = buf.getvalue()
res = {"image": "4000x3000 synthetic photograph-like JPEG q90 4:2:0", "jpeg_MB": round(len(data) / 1e6, 2)}
t0 = time.perf_counter()
for _ in range(4):
    np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
res["pil_images_per_s_1thread"] = round(4 / (time.perf_counter() - t0), 1)
for exact in (True, False):
  tag = "exact" if exact else "nvjpeg"
  for nt in (1, 4, 8, 16):
    pool = DecodePool(nt, exact=exact)
    pool.decode_many([data] * nt)
    torch.cuda.synchronize()
    n = 8 * nt
    t0 = time.perf_counter()
    out = pool.decode_many([data] * n)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert all(e is None for _, e in out)
    res[f"{tag}_images_per_s_{nt}threads"] = round(n / dt, 1)
    pool.close()
print(json.dumps(res))
