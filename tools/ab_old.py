"""A/B helper: run bench.py against another build of the library (MC_LIB=<path to .so>), same box, same process setup."""
import os, runpy, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mermaid_classifier_b200 import _lib, build, inference
if os.environ.get("MC_LIB"):
    _lib.LIB_PATH = Path(os.environ["MC_LIB"])
    build.is_stale = lambda: False
    if "old" in os.environ["MC_LIB"]:
        _lib.SIGNATURES.pop("mc_head_set_exact", None)
        inference.DeviceHead._set_exact = lambda self, e: None
sys.argv = ["bench.py"] + sys.argv[1:]
runpy.run_path(str(Path(__file__).resolve().parents[1] / "bench.py"), run_name="__main__")
