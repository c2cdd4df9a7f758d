"""Per-frame uploads from page-locked memory run ahead of the frame that uses them (tr_api.cu upload_ahead: own stream,
ping-pong buffers, released by the enqueued frame's begin mark).  A frame must see exactly the instances and lights uploaded
before it — never the previous frame's, never the next one's — whatever the host does to the source afterwards."""
import numpy as np
import pytest

from transmission_renderer_b200 import Renderer, abi, host, scenes

pytestmark = pytest.mark.gpu


def _upload_static(r, lut, s):
    r.set_uniforms(s["uniforms"])
    r.set_materials(s["materials"])
    r.set_ggx_lut(lut)
    r.set_primitives(s["primitives"])
    m = s["mesh"]
    r.set_mesh(m["positions"], m["normals"], m["uvs"], m["indices"])
    r.build_clusters(s["camera"].write_cluster_data())


def _variants(s, n):
    """n versions of the scene's instances (shifted sideways) and lights (scaled emission)."""
    out = []
    for k in range(n):
        inst = s["instances"].copy()
        inst["translation_and_scale"][:, 0] += 0.35 * k
        lights = s["lights"].copy()
        lights["colour_emission_and_falloff_distance_sq"][:, :3] *= 1.0 + 0.5 * k
        out.append((inst, lights))
    return out


def test_frames_see_their_own_uploads(ggx_lut):
    import torch
    w, h = 480, 270
    s = scenes.instanced_scene(w, h, n_instances=3000, n_lights=32)
    s["instances"] = np.ascontiguousarray(s["instances"]).astype(abi.instance)
    s["lights"] = np.ascontiguousarray(s["lights"]).astype(abi.light)
    fp = s["camera"].frame_params(host.default_tonemap_params())
    versions = _variants(s, 4)

    # reference: one fresh context per version, pageable sources (the copy goes into the compute stream)
    want = []
    for inst, lights in versions:
        with Renderer(w, h) as r:
            _upload_static(r, ggx_lut, s)
            r.set_instances(inst)
            r.set_lights(lights)
            r.frame(fp)
            want.append((r.read_hdr().copy(), r.read_visible_instances().copy()))
    assert any((want[0][0] != want[k][0]).any() for k in range(1, 4)), "the versions must differ for the test to mean anything"

    # one context, page-locked sources rewritten in place right after every enqueue, frames back to back without a sync
    inst_pin = torch.empty(versions[0][0].nbytes, dtype=torch.uint8).pin_memory()
    lights_pin = torch.empty(versions[0][1].nbytes, dtype=torch.uint8).pin_memory()
    inst_host = inst_pin.numpy().view(abi.instance)
    lights_host = lights_pin.numpy().view(abi.light)
    with Renderer(w, h) as r:
        _upload_static(r, ggx_lut, s)
        got = []
        for rep in range(2):                      # the second round starts with both ping-pong buffers in use
            for k, (inst, lights) in enumerate(versions):
                inst_host[:] = inst
                lights_host[:] = lights
                r.set_instances(inst_host)
                r.set_lights(lights_host)
                r.frame(fp)
                r.sync()                          # the contract: a page-locked source stays unchanged until the stream has read it
                got.append((r.read_hdr().copy(), r.read_visible_instances().copy()))
        for j, (hdr, vis) in enumerate(got):
            np.testing.assert_array_equal(vis, want[j % 4][1], err_msg=f"frame {j}: visible set of another upload")
            assert hdr.tobytes() == want[j % 4][0].tobytes(), f"frame {j} was rendered from another frame's instances or lights"

        # without a sync between the frames: sources alternate between two page-locked pairs, as a double-buffered host does
        pins = []
        for _ in range(2):
            a = torch.empty(versions[0][0].nbytes, dtype=torch.uint8).pin_memory()
            b = torch.empty(versions[0][1].nbytes, dtype=torch.uint8).pin_memory()
            pins.append((a, b, a.numpy().view(abi.instance), b.numpy().view(abi.light)))
        for k in (1, 2):
            _, _, ih, lh = pins[k & 1]
            ih[:] = versions[k][0]
            lh[:] = versions[k][1]
            r.set_instances(ih)
            r.set_lights(lh)
            r.frame(fp)
        r.sync()
        assert r.read_hdr().tobytes() == want[2][0].tobytes()
        # two uploads of the same buffer before one frame: the later one wins
        inst_host[:] = versions[3][0]
        r.set_instances(inst_host)
        r.sync()
        inst_host[:] = versions[1][0]
        r.set_instances(inst_host)
        lights_host[:] = versions[1][1]
        r.set_lights(lights_host)
        r.frame(fp)
        r.sync()
        assert r.read_hdr().tobytes() == want[1][0].tobytes()
