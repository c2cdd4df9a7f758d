"""Writes tests/golden/spv_layouts.json: the explicit memory layouts (OpMemberDecorate Offset, OpDecorate ArrayStride)
of every buffer / push-constant block in the reference's shipped SPIR-V modules
(/root/reference/compiled-shaders/normal/*.spv), parsed by oracle/spv2c.py.  tests/test_abi_layout.py checks
include/tr_abi.h and transmission_renderer_b200/abi.py against it.  Run in the build container (the GPU box has no
/root/reference):  python tests/golden/make_spv_layouts.py
"""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import spv2c  # noqa: E402

out = {}
for path in sorted(glob.glob("/root/reference/compiled-shaders/normal/*.spv")):
    out[os.path.basename(path)[:-4]] = spv2c.layouts(path)
with open(os.path.join(ROOT, "tests", "golden", "spv_layouts.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print("modules:", ", ".join(out))
