"""examples/tr_sys.rs — the complete Rust binding of the C ABI for the reference (a Rust application), generated from
include/tr_abi.h by tools/gen_rust_binding.py.  No Rust toolchain exists here, so the file is checked, not compiled: it is in
sync with the header, binds every symbol libtr.so exports, and its ABI-own structs have the sizes the header asserts."""
import os
import re
import subprocess
import sys

import transmission_renderer_b200 as trb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RS = os.path.join(ROOT, "examples", "tr_sys.rs")

SIZES = {"u8": (1, 1), "u16": (2, 2), "u32": (4, 4), "i32": (4, 4), "f32": (4, 4), "u64": (8, 8), "usize": (8, 8), "[f32; 3]": (12, 4),
         "shared_structs::CullingPushConstants": (96, 16), "shared_structs::AssignLightsPushConstants": (80, 16),
         "shared_structs::PushConstants": (96, 16), "colstodian::tonemap::BakedLottesTonemapperParams": (28, 4)}


def test_binding_is_in_sync_with_the_header():
    assert subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_binding.py"), "--check"]).returncode == 0, \
        "examples/tr_sys.rs is stale: run python tools/gen_rust_binding.py"


def test_binding_covers_every_export():
    text = open(RS).read()
    bound = set(re.findall(r"pub fn (tr_\w+)\(", text))
    assert bound == set(trb.EXPORTS), (sorted(set(trb.EXPORTS) - bound), sorted(bound - set(trb.EXPORTS)))
    lib = trb.lib()
    assert all(hasattr(lib, n) for n in bound)
    # every status-returning export returns i32, the two string getters a C string
    for name in bound:
        line = next(l for l in text.splitlines() if l.strip().startswith(f"pub fn {name}("))
        ret = line[line.rindex(")") + 1:]
        assert ret == (" -> *const std::os::raw::c_char;" if name in ("tr_last_error", "tr_version") else " -> i32;"), (name, ret)


def _layout(fields, structs):
    off, align = 0, 1
    for _, t in fields:
        if t.startswith("*"):
            size, a = 8, 8
        elif t in SIZES:
            size, a = SIZES[t]
        else:
            size, a = _layout(structs[t], structs)
        off = (off + a - 1) // a * a + size
        align = max(align, a)
    return (off + align - 1) // align * align, align


def test_repr_c_structs_have_the_sizes_the_header_asserts():
    text = open(RS).read()
    structs = {m.group(1): [tuple(f.strip().removeprefix("pub ").split(": ", 1)) for f in m.group(2).strip().rstrip(",").split(",\n")]
               for m in re.finditer(r"pub struct (\w+) \{\n(.*?)\n\}", text, flags=re.S)}
    want = {"TrMaterialParams": 40, "TrBasicBrdfParams": 88, "TrBrdfResult": 24, "TrTransmissionBtdfParams": 76,
            "TrIblVolumeRefractionParams": 104, "TrPointLightParams": 100, "TrPointLightResult": 36, "TrConfig": 24,
            "TrGbufferPlanes": 64, "TrGbufferPlanesOut": 64, "TrFrameParams": 304, "TrFrameTimes": 36}
    assert set(want) <= set(structs)
    for name, size in want.items():
        assert _layout(structs[name], structs)[0] == size, (name, _layout(structs[name], structs))
