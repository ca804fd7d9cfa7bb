"""Window upload (csrc/host_pipe.inl) against the whole-image one-call path: per-patch mismatch report."""
import numpy as np, torch
from mermaid_classifier_b200 import synth
from mermaid_classifier_b200.extractor import EfficientNetExtractor

sd = synth.synth_backbone_state_dict()
ext = EfficientNetExtractor(state_dict=sd, mode="fp32", max_batch=24)
shapes = [(1500, 2000), (100, 3000), (150, 2600), (3000, 100), (900, 1200)]
counts = [12, 6, 5, 6, 40]
ims = [synth.synth_image(5, i, h, w) for i, (h, w) in enumerate(shapes)]
rcs = [synth.synth_points(5, i, h, w, c, corners=(i < 4)) for i, ((h, w), c) in enumerate(zip(shapes, counts))]
want = np.concatenate([ext.extract_array(im, rc) for im, rc in zip(ims, rcs)])
for name, srcs in (("pageable", ims), ("pinned", [torch.from_numpy(im).pin_memory() for im in ims])):
    got, _ = ext.extract_many(srcs, rcs)
    print(name, "stats", ext.pipe_stats(), "equal", np.array_equal(got, want))
    k = 0
    for i, rc in enumerate(rcs):
        for (r, c) in rc:
            d = float(np.abs(got[k] - want[k]).max())
            if d != 0:
                print("  mismatch image", i, shapes[i], "point", (r, c), "max-abs", d)
            k += 1
