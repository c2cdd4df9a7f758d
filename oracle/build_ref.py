"""oracle/build_ref.py — TEST INFRASTRUCTURE ONLY.

Builds oracle/_ref/libspvref.so: the reference's shipped SPIR-V modules (/root/reference/compiled-shaders/normal/*.spv and
the two fragment modules of compiled-shaders/ray-tracing/) translated to C by oracle/spv2c.py and linked with oracle/spv_harness.c (+ liboracle.so for the samplers the Vulkan
implementation supplied).  Everything generated goes to oracle/_ref/ (git-ignored; it travels to the GPU box with the
snapshot, /root/reference does not).  When /root/reference is absent the prebuilt library is used as it is.

    python oracle/build_ref.py [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
SPV_DIR = "/root/reference/compiled-shaders/normal"
MODULES = ["frustum_culling", "demultiplex_draws", "write_cluster_data", "assign_lights_to_clusters", "fragment",
           "fragment_transmission", "fragment_tonemap", "vertex_instanced", "vertex_instanced_with_scale",
           "depth_pre_pass_instanced", "depth_pre_pass_alpha_clip", "depth_pre_pass_vertex_alpha_clip"]
# the same two fragment entry points as the reference builds them for `--ray-tracing` (SPV_KHR_ray_query)
RT_DIR = "/root/reference/compiled-shaders/ray-tracing"
RT_MODULES = ["fragment", "fragment_transmission"]
CFLAGS = ["-O2", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-fvisibility=hidden",
          "-DSPV_ORACLE_MATH", "-w"]
SO = os.path.join(REF, "libspvref.so")


def available():
    return os.path.exists(SO)


def build(force=False):
    """Returns the path of libspvref.so, or None when neither the reference nor a prebuilt library is here."""
    if not os.path.isdir(SPV_DIR):
        return SO if os.path.exists(SO) else None
    for p in (HERE, os.path.dirname(HERE)):
        if p not in sys.path:
            sys.path.insert(0, p)
    import spv2c
    from oracle import pyoracle
    oracle_so = pyoracle.build()
    deps = [os.path.join(HERE, f) for f in ("spv2c.py", "spv_harness.c", "spv_ctx.h", "oracle.h", "build_ref.py")]
    deps += [os.path.join(SPV_DIR, m + ".spv") for m in MODULES] + [os.path.join(RT_DIR, m + ".spv") for m in RT_MODULES] + [oracle_so]
    if not force and os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in deps):
        return SO
    os.makedirs(REF, exist_ok=True)
    objs = []
    for src_dir, m, sym in [(SPV_DIR, m, m) for m in MODULES] + [(RT_DIR, m, "rt_" + m) for m in RT_MODULES]:
        c = os.path.join(REF, sym + ".c")
        with open(c, "w") as f:
            f.write(spv2c.translate(os.path.join(src_dir, m + ".spv"), "spv_" + sym))
        o = os.path.join(REF, sym + ".o")
        subprocess.check_call(["gcc"] + CFLAGS + ["-c", c, "-o", o])
        objs.append(o)
    subprocess.check_call(["gcc"] + CFLAGS + ["-shared", "-o", SO, os.path.join(HERE, "spv_harness.c")] + objs +
                          ["-L" + HERE, "-loracle", "-Wl,-rpath,$ORIGIN/..", "-lm"])
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
