"""CPU-side checks of the drop-in boundary: libtr.so loads, exports every symbol include/tr_abi.h
declares, and refuses to do anything without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import transmission_renderer_b200 as trb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "tr_abi.h")).read()
    return sorted(set(re.findall(r"TR_API\s+(?:const\s+)?\w+\*?\s+(tr_\w+)\s*\(", text)))


def test_header_and_export_list_agree():
    assert _declared() == sorted(trb.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = trb.lib()
    for name in _declared():
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.tr_version()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(trb.TrError) as e:
        trb.Renderer(64, 64)
    assert e.value.status == -3 and "no CUDA device" in str(e.value)  # TR_ERR_CUDA


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import, include, link or load it."""
    pkg = os.path.join(ROOT, "transmission_renderer_b200")
    banned = ("pyoracle", "liboracle", "import oracle", "from oracle", '#include "oracle', "oracle.h", "orc_")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                for b in banned:
                    # comments may NAME oracle functions they mirror (e.g. "oracle: orc_log2_spec"); code may not
                    code = "\n".join(l.split("//")[0].split("#")[0] if not l.lstrip().startswith("#include") else l
                                     for l in text.splitlines())
                    assert b not in code, (f, b)


def test_plain_c_host_compiles_and_fails_loudly_without_a_device(tmp_path):
    """examples/frame.c is a C11 host of the ABI: the header compiles as C, the library links, and without a CUDA device
    the first call fails with a message instead of falling back to anything."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    libdir = os.path.join(root, "transmission_renderer_b200")
    exe = str(tmp_path / "frame")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "frame.c"),
                           "-L", libdir, "-ltr", f"-Wl,-rpath,{libdir}", "-lm", "-o", exe])
    p = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""), timeout=120)
    assert p.returncode == 3 and "tr_create" in p.stderr and "no CPU path" in p.stderr
