"""extract_many (mc_extract_images_host) vs the per-image host call: equality, labels, and throughput from pinned /
pageable host images.   python tools/pipe_check.py [--images 40] [--mode fp32]"""
import argparse, json, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=40)
    ap.add_argument("--points", type=int, default=100)
    ap.add_argument("--mode", default="fp32")
    ap.add_argument("--batch", type=int, default=1000)
    args = ap.parse_args()
    import torch
    from mermaid_classifier_b200 import synth
    from mermaid_classifier_b200.extractor import EfficientNetExtractor, synth_image_device
    from mermaid_classifier_b200.inference import DeviceHead
    H, W = 3000, 4000
    sd = synth.synth_backbone_state_dict()
    ext = EfficientNetExtractor(state_dict=sd, mode=args.mode, max_batch=args.batch)
    w, bb, a, b, _ = synth.synth_head(1280, (200, 100), 500, seed=0)
    head = DeviceHead([x.numpy() for x in w], [x.numpy() for x in bb], a.numpy(), b.numpy())
    pool = min(args.images, 16)
    pinned = [torch.empty((H, W, 3), dtype=torch.uint8).pin_memory() for _ in range(pool)]
    for i in range(pool):
        pinned[i].copy_(synth_image_device(synth.DEFAULT_SEED, i, H, W))
    torch.cuda.synchronize()
    rcs = [synth.synth_points(synth.DEFAULT_SEED, i % pool, H, W, args.points) for i in range(args.images)]
    ims = [pinned[i % pool] for i in range(args.images)]
    # small case first: equality with the one-image call (ragged: one image without points, odd shapes)
    small = [synth.synth_image(3, i, 300 + 40 * i, 500 - 30 * i) for i in range(4)]
    srcs = [synth.synth_points(3, i, small[i].shape[0], small[i].shape[1], [7, 0, 30, 12][i], corners=(i == 0)) for i in range(4)]
    f_many, l_many = ext.extract_many(small, srcs, head=head)
    f_one = np.concatenate([ext.extract_array(im, rc) for im, rc in zip(small, srcs) if len(rc)])
    l_one = head.scores_host(f_one)[1]
    res = {"small_equal": bool(np.array_equal(f_many, f_one)), "small_labels_equal": bool(np.array_equal(l_many, l_one)),
           "small_max_abs": float(np.abs(f_many - f_one).max())}
    for name, src in (("pinned", ims), ("pageable", [t.numpy().copy() for t in ims[: min(20, len(ims))]])):
        r = rcs[: len(src)]
        ext.extract_many(src[:10], r[:10], head=head)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        f, l = ext.extract_many(src, r, head=head)
        dt = time.perf_counter() - t0
        st = ext.pipe_stats()
        res[name] = {"patches_per_s": round(f.shape[0] / dt, 1), "h2d_GBps": round(st["h2d"] / dt / 1e9, 2), "groups": st["groups"], "n": int(f.shape[0])}
    print(json.dumps(res), flush=True)

if __name__ == "__main__":
    main()
