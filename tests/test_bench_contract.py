"""bench.py contract, the parts that run without a GPU: the reference arm prints one JSON line with the keys the driver reads,
and the product arm refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, **env):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600,
                          env=dict(os.environ, CUDA_VISIBLE_DEVICES="", **env))


@pytest.mark.parametrize("code", ["spirv", "port"])
def test_reference_arm_line(code):
    """code = spirv: the reference's own compiled shader modules shade (oracle/_ref/libspvref.so; kind "reference", or "port"
    with a note where that library is absent); code = port: the C restatement is forced (TR_CPU_CODE=port)."""
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "0", TR_CPU_CODE=code)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "shaded_mpixels_per_s" and d["unit"] == "Mpx/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("configs[3]") and d["config"]["width"] == 3840 and d["config"]["lights"] == 64
    cb = d["cpu_baseline"]
    from oracle import build_ref
    kind = "reference" if code == "spirv" and build_ref.build() is not None else "port"
    assert cb["kind"] == kind and cb["cores"] >= 1 and cb["value"] == d["value"] and "rows" in cb["sample"]
    assert (cb["value_port"] is not None and cb["value_port"] > 0) == (kind == "reference")
    assert ("fragment.spv" in cb["sample"]) == (kind == "reference")
    assert d["e2e"] == {"value": d["value"], "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_needs_cuda():
    p = _run("--steps", "1", "--warmup", "0")
    assert p.returncode != 0
    assert "no CUDA device" in (p.stderr + p.stdout)
