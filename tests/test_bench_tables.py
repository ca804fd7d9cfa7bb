"""The roofline arithmetic bench.py reports (algorithmic HBM bytes per patch per layer) is pinned on the CPU: to SURVEY
section 8d's element counts (depthwise 3.788 M in + 2.307 M out; 54.97 MB fp32 / 27.60 MB bf16 per patch), to the oracle's
own layer table, and to itself under fusion (a fused kernel carries exactly the bytes of the layers it replaces)."""
import importlib.util
import sys
from pathlib import Path

import pytest

from oracle import effnet

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["bench_under_test"] = mod
    spec.loader.exec_module(mod)
    return mod


def test_block_table_is_the_oracles(bench):
    blocks = effnet.b0_blocks()
    assert len(bench.B0) == len(blocks) == 16
    h = 112
    for (k, s, ex, ci, co, h_in), b in zip(bench.B0, blocks):
        assert (k, s, ex, ci, co) == (b.kernel, b.stride, b.expand, b.c_in, b.c_out)
        assert h_in == h
        h = (h + s - 1) // s
    assert h == 7


@pytest.mark.parametrize("e,total_mb", [(4, 54.97), (2, 27.60)])
def test_totals_match_the_survey(bench, monkeypatch, e, total_mb):
    monkeypatch.setenv("MC_FUSE_MASK", "0")
    lb = bench.layer_bytes(e)
    assert abs(sum(v for _, v in lb.values()) / 1e6 - total_mb) < 0.01
    dw_in = sum(h * h * ci * ex for (_k, _s, ex, ci, _co, h) in bench.B0)
    dw_out = sum(((h + s - 1) // s) ** 2 * ci * ex for (_k, s, ex, ci, _co, h) in bench.B0)
    assert round(dw_in / 1e6, 3) == 3.788 and round(dw_out / 1e6, 3) == 2.307
    assert sum(v for n, v in lb.values() if n.endswith(".depthwise")) == (dw_in + dw_out) * e
    assert lb[0][1] == 224 * 224 * 3 + 112 * 112 * 32 * e


@pytest.mark.parametrize("e", [4, 2])
@pytest.mark.parametrize("mask", ["2", "6", "e"])
def test_fusion_conserves_algorithmic_bytes(bench, monkeypatch, e, mask):
    monkeypatch.setenv("MC_FUSE_MASK", "0")
    plain = bench.layer_bytes(e)
    monkeypatch.setenv("MC_FUSE_MASK", mask)
    fused = bench.layer_bytes(e)
    assert sum(v for _, v in fused.values()) == sum(v for _, v in plain.values())
    for b in range(16):
        if (int(mask, 16) >> b) & 1:
            assert 1 + 4 * b not in fused and "fused" in fused[2 + 4 * b][0]
            assert fused[2 + 4 * b][1] == plain[1 + 4 * b][1] + plain[2 + 4 * b][1]
        else:
            assert fused.get(1 + 4 * b) == plain.get(1 + 4 * b) and fused[2 + 4 * b] == plain[2 + 4 * b]


def test_defaults_follow_the_library(bench, monkeypatch):
    """bench.fused_blocks mirrors csrc/api.cu's MC_FUSE_DEFAULT_FP32 / MC_FUSE_DEFAULT."""
    monkeypatch.delenv("MC_FUSE_MASK", raising=False)
    src = (ROOT / "mermaid_classifier_b200" / "csrc" / "api.cu").read_text()
    import re

    d16 = int(re.search(r"#define MC_FUSE_DEFAULT (0x[0-9A-Fa-f]+)u", src).group(1), 16)
    d32 = int(re.search(r"#define MC_FUSE_DEFAULT_FP32 (0x[0-9A-Fa-f]+)u", src).group(1), 16)
    assert bench.fused_blocks(4) == d32 and bench.fused_blocks(2) == d16


def test_reference_arm_contract():
    """`bench.py --impl reference`: rank 0 alone runs the CPU path and prints ONE JSON line with the arm's keys; any other rank
    exits 0 without output (the driver launches the arm under torchrun like the B200 arm)."""
    import json
    import os
    import subprocess

    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1", MASTER_ADDR="127.0.0.1", MASTER_PORT="29871")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--points", "6"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "point-patches/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "point-patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1
