"""Dev-time tool: measure the pre-BatchNorm scalar mean/variance of every BN layer of the
synthetic EfficientNet-B0 (layer by layer, each layer seeing the already-calibrated
layers before it) and rewrite the ``_BN_CALIB`` table in
``mermaid_classifier_b200/synth.py``.  Uses the CPU oracle forward; run once, commit the table.

    python tools/calibrate_synth.py
"""
import re
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from mermaid_classifier_b200 import synth  # noqa: E402
from oracle import crop, effnet  # noqa: E402


def main():
    im = synth.synth_image(synth.DEFAULT_SEED, 0, 1200, 1600)
    pts = synth.synth_points(synth.DEFAULT_SEED, 0, 1200, 1600, 32, corners=True)
    x = torch.from_numpy(crop.normalize_patches(crop.crop_patches(im, pts)))
    calib: dict[str, tuple[float, float]] = {}
    orig_bn = effnet.bn

    order = ["_bn0"]
    for cfg in effnet.b0_blocks():
        p = f"_blocks.{cfg.index}."
        if cfg.expand != 1:
            order.append(p + "_bn0")
        order += [p + "_bn1", p + "_bn2"]
    order.append("_bn1")

    for target in order:
        sd = effnet.strip_module_prefix(synth.synth_backbone_state_dict(calib=calib))
        seen = {}

        def spy(xx, sdd, prefix):
            if prefix == target and prefix not in seen:
                seen[prefix] = (float(xx.mean()), float(xx.var()))
                raise StopIteration
            return orig_bn(xx, sdd, prefix)

        effnet.bn = spy
        try:
            effnet.extract_features(sd, x)
        except StopIteration:
            pass
        finally:
            effnet.bn = orig_bn
        calib[target] = seen[target]
        print(target, "mean %.4g var %.4g" % seen[target], flush=True)

    body = "_BN_CALIB: dict = {\n" + "".join(
        f'    "{k}": ({m:.6g}, {v:.6g}),\n' for k, (m, v) in calib.items()
    ) + "}\n"
    path = ROOT / "mermaid_classifier_b200" / "synth.py"
    src = path.read_text()
    src = re.sub(r"(# BEGIN _BN_CALIB[^\n]*\n).*?(# END _BN_CALIB)", lambda m: m.group(1) + body + m.group(2), src, flags=re.S)
    path.write_text(src)
    print("table written")


if __name__ == "__main__":
    main()
