"""Cross-check the EfficientNet-B0 oracle against an independent construction of the same
topology: torchvision's ``efficientnet_b0`` with stride-2 convs re-padded TF-"SAME"
(asymmetric) and every BatchNorm at eps=1e-3, loaded through a pyspacer->torchvision key map."""
import pytest
import torch

from mermaid_classifier_b200 import synth
from oracle import crop, effnet


def _tv_model(sd):
    tv = pytest.importorskip("torchvision")
    net = tv.models.efficientnet_b0(weights=None).eval()
    sd = effnet.strip_module_prefix(sd)

    def load_cna(cna, conv_key, bn_key):
        conv, bnm = cna[0], cna[1]
        conv.weight.data.copy_(sd[conv_key + ".weight"])
        bnm.weight.data.copy_(sd[bn_key + ".weight"])
        bnm.bias.data.copy_(sd[bn_key + ".bias"])
        bnm.running_mean.data.copy_(sd[bn_key + ".running_mean"])
        bnm.running_var.data.copy_(sd[bn_key + ".running_var"])

    load_cna(net.features[0], "_conv_stem", "_bn0")
    idx = 0
    for stage in list(net.features)[1:8]:
        for mb in stage:
            p = f"_blocks.{idx}."
            layers = list(mb.block)
            j = 0
            if len(layers) == 4:
                load_cna(layers[0], p + "_expand_conv", p + "_bn0")
                j = 1
            load_cna(layers[j], p + "_depthwise_conv", p + "_bn1")
            se = layers[j + 1]
            se.fc1.weight.data.copy_(sd[p + "_se_reduce.weight"])
            se.fc1.bias.data.copy_(sd[p + "_se_reduce.bias"])
            se.fc2.weight.data.copy_(sd[p + "_se_expand.weight"])
            se.fc2.bias.data.copy_(sd[p + "_se_expand.bias"])
            load_cna(layers[j + 2], p + "_project_conv", p + "_bn2")
            idx += 1
    assert idx == 16
    load_cna(net.features[8], "_conv_head", "_bn1")
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eps = 1e-3
    # TF-SAME asymmetric padding on the stride-2 convs: k3 -> (0,1), k5 -> (1,2)
    def repad(parent, name):
        conv = getattr(parent, name)
        k = conv.kernel_size[0]
        pad = (0, 1, 0, 1) if k == 3 else (1, 2, 1, 2)
        conv.padding = (0, 0)
        setattr(parent, name, torch.nn.Sequential(torch.nn.ZeroPad2d(pad), conv))

    for cna in [m for m in net.modules() if type(m).__name__ == "Conv2dNormActivation"]:
        if cna[0].stride == (2, 2):
            repad(cna, "0")
    return net


def test_layer_table_counts():
    blocks = effnet.b0_blocks()
    assert len(blocks) == 16
    assert [b.c_se for b in blocks] == [8, 4, 6, 6, 10, 10, 20, 20, 20, 28, 28, 28, 48, 48, 48, 48]
    assert [b.has_skip for b in blocks] == [False, False, True, False, True, False, True, True, False, True, True,
                                            False, True, True, True, False]
    assert effnet.same_pad(224, 3, 2) == (0, 1)
    assert effnet.same_pad(56, 5, 2) == (1, 2)
    assert effnet.same_pad(14, 5, 2) == (1, 2)
    assert effnet.same_pad(28, 3, 2) == (0, 1)
    assert effnet.same_pad(112, 3, 1) == (1, 1)
    assert effnet.same_pad(14, 5, 1) == (2, 2)


def test_state_dict_layout(backbone_sd):
    sd = effnet.strip_module_prefix(backbone_sd)
    assert all(k.startswith("module.") for k in backbone_sd)
    assert "_blocks.0._expand_conv.weight" not in sd and "_blocks.1._expand_conv.weight" in sd
    n_params = sum(v.numel() for k, v in sd.items() if not k.startswith("_fc") and "num_batches" not in k
                   and "running" not in k)
    assert n_params == 4007548  # torchvision efficientnet_b0 feature-extractor parameter count


def test_oracle_matches_torchvision_topology(backbone_sd):
    net = _tv_model(backbone_sd)
    im = synth.synth_image(synth.DEFAULT_SEED, 1, 400, 520)
    pts = synth.synth_points(synth.DEFAULT_SEED, 1, 400, 520, 6, corners=True)
    x = torch.from_numpy(crop.normalize_patches(crop.crop_patches(im, pts)))
    with torch.no_grad():
        want = torch.flatten(net.avgpool(net.features(x)), 1)
    got = effnet.extract_features(backbone_sd, x)
    assert got.shape == (len(pts), 1280)
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-4), (got - want).abs().max()
    assert 0.05 < got.std().item() < 2.0  # synthetic weights keep features O(0.1-1)


def test_batched_equals_unbatched(backbone_sd):
    x = torch.from_numpy(crop.normalize_patches(
        crop.crop_patches(synth.synth_image(1, 2, 300, 300), [(5, 5), (150, 150), (299, 0)])))
    a = effnet.extract_features(backbone_sd, x)
    b = effnet.extract_features_batched(backbone_sd, x, batch_size=2)
    assert torch.allclose(a, b, atol=1e-5)


def _hf_model(sd):
    """A second, independent construction: the Hugging Face ``transformers`` port of the Keras EfficientNet
    (``modeling_efficientnet.py``), which carries the TF-"SAME" asymmetric padding natively (ZeroPad2d + valid convs
    on the stride-2 layers), built from a B0 config with random init and loaded through a pyspacer -> HF key map."""
    tr = pytest.importorskip("transformers")
    cfg = tr.EfficientNetConfig(width_coefficient=1.0, depth_coefficient=1.0, image_size=224, hidden_dim=1280,
                                batch_norm_eps=1e-3, hidden_act="swish")
    net = tr.EfficientNetModel(cfg).eval()
    sd = effnet.strip_module_prefix(sd)
    out = {}

    def conv(dst, src):
        out[dst + ".weight"] = sd[src + ".weight"]
        if src + ".bias" in sd:
            out[dst + ".bias"] = sd[src + ".bias"]

    def bn(dst, src):
        for f in ("weight", "bias", "running_mean", "running_var"):
            out[f"{dst}.{f}"] = sd[f"{src}.{f}"]

    conv("embeddings.convolution", "_conv_stem")
    bn("embeddings.batchnorm", "_bn0")
    for i, blk in enumerate(effnet.b0_blocks()):
        p, q = f"_blocks.{i}.", f"encoder.blocks.{i}."
        if blk.expand != 1:
            conv(q + "expansion.expand_conv", p + "_expand_conv")
            bn(q + "expansion.expand_bn", p + "_bn0")
        conv(q + "depthwise_conv.depthwise_conv", p + "_depthwise_conv")
        bn(q + "depthwise_conv.depthwise_norm", p + "_bn1")
        conv(q + "squeeze_excite.reduce", p + "_se_reduce")
        conv(q + "squeeze_excite.expand", p + "_se_expand")
        conv(q + "projection.project_conv", p + "_project_conv")
        bn(q + "projection.project_bn", p + "_bn2")
    conv("encoder.top_conv", "_conv_head")
    bn("encoder.top_bn", "_bn1")
    missing, unexpected = net.load_state_dict(out, strict=False)
    assert not unexpected and all(k.endswith("num_batches_tracked") for k in missing), (missing, unexpected)
    return net


def test_oracle_matches_transformers_keras_port(backbone_sd):
    net = _hf_model(backbone_sd)
    im = synth.synth_image(synth.DEFAULT_SEED, 2, 420, 380)
    pts = synth.synth_points(synth.DEFAULT_SEED, 2, 420, 380, 6, corners=True)
    x = torch.from_numpy(crop.normalize_patches(crop.crop_patches(im, pts)))
    with torch.no_grad():
        want = net(pixel_values=x).pooler_output
    got = effnet.extract_features(backbone_sd, x)
    assert got.shape == want.shape == (len(pts), 1280)
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-4), (got - want).abs().max()
