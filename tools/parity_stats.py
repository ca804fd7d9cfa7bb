"""Error statistics of the extraction path against the CPU oracle on N patches, for a few library configurations.
   python tools/parity_stats.py [--images 20] [--mode fp32]"""
import argparse, json, os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=20)
    ap.add_argument("--points", type=int, default=100)
    ap.add_argument("--mode", default="fp32")
    args = ap.parse_args()
    import torch
    from mermaid_classifier_b200 import synth
    from mermaid_classifier_b200.extractor import EfficientNetExtractor
    from oracle import crop as ocrop, effnet as oeff
    H, W = 1200, 1600
    sd = synth.synth_backbone_state_dict()
    ims = [synth.synth_image(synth.DEFAULT_SEED, 500 + i, H, W) for i in range(args.images)]
    rcs = [synth.synth_points(synth.DEFAULT_SEED, 500 + i, H, W, args.points) for i in range(args.images)]
    want = np.concatenate([oeff.extract_features_batched(sd, torch.from_numpy(ocrop.normalize_patches(ocrop.crop_patches(im, rc))), 100).numpy()
                           for im, rc in zip(ims, rcs)])
    # the oracle against itself in float64 accumulation is not available; report its scale instead
    for name, env in (("default", {}), ("no_fuse", {"MC_FUSE_MASK": "0"}), ("no_pool_fusion", {"MC_NO_POOL_FUSION": "1"}),
                      ("simt_gemm", {"MC_TC_MASK": "0", "MC_FUSE_MASK": "0"}),
                      ("hi_written", {"MC_TC_EXP": "8", "MC_FUSE_MASK": "0"}),
                      ("expand_simt", {"MC_TC_MASK": "aaaaaaaa", "MC_FUSE_MASK": "0"}),
                      ("project_simt", {"MC_TC_MASK": "155555555", "MC_FUSE_MASK": "0"})):
        for k in ("MC_FUSE_MASK", "MC_NO_POOL_FUSION", "MC_TC_MASK", "MC_TC_EXP"):
            os.environ.pop(k, None)
        os.environ.update(env)
        ext = EfficientNetExtractor(state_dict=sd, mode=args.mode, max_batch=1000)
        got, _ = ext.extract_many(ims, rcs)
        ext.close()
        d = np.abs(got - want)
        i = np.unravel_index(d.argmax(), d.shape)
        cos = (got * want).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(want, axis=1))
        print(json.dumps({"config": name, "n": int(got.shape[0]), "max_abs": float(d.max()), "at_value": float(want[i]),
                          "feat_abs_max": float(np.abs(want).max()), "p99.99_abs": float(np.quantile(d, 0.9999)),
                          "mean_abs": float(d.mean()), "max_rel_to_rowmax": float((d.max(1) / np.abs(want).max(1)).max()),
                          "min_cos": float(cos.min())}), flush=True)

if __name__ == "__main__":
    main()
