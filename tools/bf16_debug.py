import os, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mermaid_classifier_b200 import synth
from mermaid_classifier_b200.extractor import EfficientNetExtractor
from oracle import crop as ocrop, effnet as oeff
mode = sys.argv[1]; nb = int(sys.argv[2])
sd = synth.synth_backbone_state_dict()
im = synth.synth_image(synth.DEFAULT_SEED, 5, 800, 800)
pts = synth.synth_points(synth.DEFAULT_SEED, 5, 800, 800, nb)
ext = EfficientNetExtractor(state_dict=sd, mode=mode, max_batch=nb)
got = ext.extract_array(im, pts)
torch.cuda.synchronize()
k = min(nb, 12)
want = oeff.extract_features(sd, torch.from_numpy(ocrop.normalize_patches(ocrop.crop_patches(im, pts[:k])))).numpy()
g = got[:k]
cos = (g * want).sum(1) / (np.linalg.norm(g, axis=1) * np.linalg.norm(want, axis=1))
print(mode, nb, 'mask', os.environ.get('MC_TC_MASK'), 'min cos', cos.min(), 'max abs', np.abs(g - want).max(), flush=True)
