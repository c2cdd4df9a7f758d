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
