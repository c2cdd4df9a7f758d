"""Row N4: glTF ingest (transmission_renderer_b200/gltf_ingest.py) against the semantics of the reference's loader
(src/model_loading.rs).  glTF-Sample-Models is not available offline, so the asset is written here with write_gltf."""
import json
import math
import os

import numpy as np
import pytest

from transmission_renderer_b200 import abi, gltf_ingest, host, scenes

f32 = np.float32


def _asset(tmp_path):
    sphere = scenes.uv_sphere(16, 8)
    knot = scenes.torus_knot(n_u=48, n_v=8)
    quad = scenes.quad_mesh(1.0)
    tex = scenes.procedural_textures(32)
    images = [t["levels"][0] for t in tex[:3]]                       # diffuse-like, metal-rough-like, normal-map-like
    materials = [
        dict(pbrMetallicRoughness=dict(baseColorFactor=[0.9, 0.8, 0.7, 1.0], metallicFactor=0.0, roughnessFactor=0.4,
                                       baseColorTexture=dict(index=0, extensions=dict(KHR_texture_transform=dict(scale=[2.0, 3.0]))),
                                       metallicRoughnessTexture=dict(index=0)),
             normalTexture=dict(index=2, scale=0.8), emissiveFactor=[0.1, 0.2, 0.3],
             extensions=dict(KHR_materials_specular=dict(specularFactor=0.7, specularTexture=dict(index=0)))),
        dict(alphaMode="MASK", alphaCutoff=0.7, pbrMetallicRoughness=dict(baseColorTexture=dict(index=0))),
        dict(pbrMetallicRoughness=dict(roughnessFactor=0.25, metallicFactor=0.0),
             extensions=dict(KHR_materials_transmission=dict(transmissionFactor=0.9),
                             KHR_materials_volume=dict(thicknessFactor=0.5, attenuationDistance=0.6, attenuationColor=[0.9, 0.4, 0.2]),
                             KHR_materials_ior=dict(ior=1.33))),
        dict(alphaMode="MASK", extensions=dict(KHR_materials_transmission=dict())),
        dict(),                                                       # all defaults
    ]

    def prim(mesh, material):
        return dict(positions=mesh[0], normals=mesh[1], uvs=mesh[2], indices=mesh[3], material=material)

    def moved(mesh, offset):                                         # primitives of one glTF mesh share the node, not the place
        return (mesh[0] + np.asarray(offset, f32),) + tuple(mesh[1:])

    meshes = [[prim(sphere, 0), prim(moved(sphere, (2.4, 0.0, 0.0)), 1)],
              [prim(knot, 2), dict(prim(moved(knot, (-1.6, 0.0, 0.4)), 3), uvs=None)], [prim(quad, 4)]]
    c, s_ = math.cos(0.35), math.sin(0.35)
    matrix = np.array([[1.5 * c, 0, 1.5 * s_, 0.4], [0, 1.5, 0, 2.0], [-1.5 * s_, 0, 1.5 * c, 1.5], [0, 0, 0, 1]], np.float64)
    nodes = [
        dict(translation=[0.5, 1.0, -2.0], rotation=[0.0, math.sin(0.3), 0.0, math.cos(0.3)], scale=[2.0, 2.0, 2.0], children=[1]),
        dict(mesh=0, translation=[0.2, 0.0, 0.1], rotation=[math.sin(0.2), 0.0, 0.0, math.cos(0.2)], scale=[0.5, 0.5, 0.5]),
        dict(mesh=1, matrix=[float(x) for x in matrix.T.reshape(-1)]),
        dict(mesh=2, translation=[0.0, 0.0, -4.0], scale=[30.0, 30.0, 30.0]),
    ]
    path = gltf_ingest.write_gltf(str(tmp_path / "asset.gltf"), nodes, meshes, materials, images)
    return path, dict(sphere=sphere, knot=knot, quad=quad, matrix=matrix, images=images)


def test_ingest_follows_the_reference_loader(tmp_path):
    path, src = _asset(tmp_path)
    m = gltf_ingest.finish(gltf_ingest.load_gltf(path))
    prims, inst, mats = m["primitives"], m["instances"], m["materials"]
    # one instance and one PrimitiveInfo per glTF primitive, in node order (model_loading.rs:58-160)
    assert len(prims) == len(inst) == 5 and list(inst["primitive_id"]) == [0, 1, 2, 3, 4] and list(inst["material_id"]) == [0, 1, 2, 3, 4]
    assert list(prims["first_instance"]) == [0, 1, 2, 3, 4]
    # draw buffer from alpha mode x transmission (:68-78) and the draw-count bookkeeping (:80-85)
    assert list(prims["draw_buffer_index"]) == [0, 1, 2, 3, 0]
    assert m["max_draw_counts"] == dict(opaque=2, alpha_clip=1, transmission=1, transmission_alpha_clip=1)
    # index / vertex layout: indices rebased onto the shared vertex arrays
    n_s, n_k, n_q = len(src["sphere"][0]), len(src["knot"][0]), len(src["quad"][0])
    assert len(m["mesh"]["positions"]) == 2 * n_s + 2 * n_k + n_q
    assert list(prims["first_index"]) == list(np.cumsum([0] + [len(src[k][3]) for k in ("sphere", "sphere", "knot", "knot")]))
    i1 = m["mesh"]["indices"][prims["first_index"][1]:prims["first_index"][1] + prims["index_count"][1]]
    assert n_s <= i1.min() and i1.max() < 2 * n_s and np.array_equal(i1 - n_s, src["sphere"][3])
    # bounding sphere = bbox centre / half diagonal (:148-155)
    lo, hi = src["sphere"][0].min(0), src["sphere"][0].max(0)
    np.testing.assert_allclose(prims["packed_bounding_sphere"][0], [*((lo + hi) / 2), np.linalg.norm(hi - lo) / 2], rtol=1e-6, atol=1e-7)
    # node hierarchy: parent * child, as matrices (NodeTree::transform_of, :471-483)
    def mat(t, q, s):
        x, y, z, w = q
        r = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        out = np.eye(4)
        out[:3, :3] = r * s
        out[:3, 3] = t
        return out
    parent = mat([0.5, 1.0, -2.0], [0.0, math.sin(0.3), 0.0, math.cos(0.3)], 2.0)
    child = mat([0.2, 0.0, 0.1], [math.sin(0.2), 0.0, 0.0, math.cos(0.2)], 0.5)
    for k, expect in ((0, parent @ child), (2, src["matrix"])):
        got = mat(inst["translation_and_scale"][k, :3], inst["rotation"][k], inst["translation_and_scale"][k, 3])
        np.testing.assert_allclose(got, expect, atol=2e-6)
    assert abs(inst["translation_and_scale"][0, 3] - 1.0) < 1e-6 and abs(inst["translation_and_scale"][2, 3] - 1.5) < 1e-6
    # uv scaling of the base colour texture transform (:87-95), zero uvs when a primitive has none (:121-134)
    uv0 = m["mesh"]["uvs"][:n_s]
    np.testing.assert_allclose(uv0, src["sphere"][2] * np.array([2.0, 3.0], f32), rtol=1e-6)
    assert (m["mesh"]["uvs"][2 * n_s + n_k:2 * n_s + 2 * n_k] == 0).all()
    # materials (:231-333)
    assert abs(mats["roughness_factor"][0] - 0.4) < 1e-7 and abs(mats["normal_map_scale"][0] - 0.8) < 1e-7
    assert abs(mats["alpha_clipping_cutoff"][1] - 0.7) < 1e-7 and mats["alpha_clipping_cutoff"][0] == 0.5
    assert abs(mats["index_of_refraction"][2] - 1.33) < 1e-7 and mats["index_of_refraction"][0] == 1.5
    assert abs(mats["transmission_factor"][2] - 0.9) < 1e-7 and mats["transmission_factor"][3] == 0.0 and mats["transmission_factor"][0] == 0.0
    assert abs(mats["thickness_factor"][2] - 0.5) < 1e-7 and abs(mats["attenuation_distance"][2] - 0.6) < 1e-7
    assert np.isinf(mats["attenuation_distance"][0]) and (mats["attenuation_colour"][4, :3] == 1).all()
    np.testing.assert_allclose(mats["attenuation_colour"][2, :3], [0.9, 0.4, 0.2], rtol=1e-6)
    assert abs(mats["specular_factor"][0] - 0.7) < 1e-7 and mats["specular_factor"][4] == 1.0 and mats["normal_map_scale"][4] == 0.0
    # images: one binding per (image, sRGB-ness); the specular texture (alpha only) reuses the sRGB binding (:176-187)
    t0 = mats["textures"][0]
    slot = scenes.TEX_SLOTS
    assert t0[slot["diffuse"]] != t0[slot["metallic_roughness"]]                   # same image, sRGB and linear
    assert t0[slot["specular"]] == t0[slot["diffuse"]]
    assert mats["textures"][1][slot["diffuse"]] == t0[slot["diffuse"]]             # cached
    assert m["textures"][t0[slot["diffuse"]]]["srgb"] and not m["textures"][t0[slot["metallic_roughness"]]]["srgb"]
    assert (mats["textures"][4] == -1).all()
    assert len(m["textures"][0]["levels"]) == host.mip_levels_for_size(32, 32)


def test_base_transform_and_roughness_override(tmp_path):
    """`--scale` and `--roughness-override` of the reference CLI (src/main.rs:65-91): the base Similarity multiplies every
    node transform and the attenuation distance (model_loading.rs:62, 317); the override replaces roughness (:294)."""
    path, _ = _asset(tmp_path)
    base = gltf_ingest.Similarity((1.0, 0.0, 0.0), (0, 0, 0, 1), 3.0)
    a = gltf_ingest.finish(gltf_ingest.load_gltf(path))
    b = gltf_ingest.finish(gltf_ingest.load_gltf(path, base_transform=base, roughness_override=0.15))
    np.testing.assert_allclose(b["instances"]["translation_and_scale"][:, 3], 3.0 * a["instances"]["translation_and_scale"][:, 3], rtol=1e-6)
    np.testing.assert_allclose(b["instances"]["translation_and_scale"][:, :3], 3.0 * a["instances"]["translation_and_scale"][:, :3] + [1, 0, 0], rtol=1e-5, atol=1e-5)
    assert abs(b["materials"]["attenuation_distance"][2] - 1.8) < 1e-6 and (np.abs(b["materials"]["roughness_factor"] - 0.15) < 1e-7).all()


def test_two_models_share_one_set_of_buffers(tmp_path):
    """The reference loads Sponza and then the model into the same ModelStagingBuffers (src/main.rs:342-370): ids continue."""
    path, _ = _asset(tmp_path)
    m = gltf_ingest.load_gltf(path)
    m = gltf_ingest.load_gltf(path, base_transform=gltf_ingest.Similarity((5, 0, 0)), into=m)
    s = gltf_ingest.finish(m)
    assert len(s["primitives"]) == 10 and list(s["instances"]["material_id"][5:]) == [5, 6, 7, 8, 9]
    assert s["primitives"]["first_index"][5] == s["primitives"]["first_index"][4] + s["primitives"]["index_count"][4]
    assert s["mesh"]["indices"][s["primitives"]["first_index"][5]:].min() >= len(s["mesh"]["positions"]) // 2


def test_ingested_scene_renders_on_the_oracle(oracle, ggx_lut, tmp_path):
    path, _ = _asset(tmp_path)
    s = gltf_ingest.finish(gltf_ingest.load_gltf(path))
    w, h = 160, 90
    cam = scenes.Camera(w, h, (0.0, 3.0, 6.0), 0.0, -15.0)
    _, visible = oracle.frustum_culling(s["instances"], s["primitives"], cam.culling())
    g0, g1 = oracle.visibility(s["mesh"], s["instances"], s["primitives"], visible, cam.push_constants(), derivatives=True,
                               materials=s["materials"], textures=s["textures"])
    assert (g0["depth"] > 0).mean() > 0.3 and (g1["depth"] > 0).mean() > 0.005
    assert set(np.unique(g0["material_id"][g0["depth"] > 0])) == {0, 1, 4} and set(np.unique(g1["material_id"][g1["depth"] > 0])) == {2, 3}
    # the MASK material really drops fragments: without the alpha test its sphere covers more pixels
    s0, _ = oracle.visibility(s["mesh"], s["instances"], s["primitives"], visible, cam.push_constants())
    assert (s0["material_id"] == 1).sum() > (g0["material_id"] == 1).sum() > 0


def _to_glb(gltf_path, glb_path):
    """Repack a .gltf + .bin + PNG files as one binary container with the images behind buffer views."""
    doc = json.load(open(gltf_path))
    base = os.path.dirname(gltf_path)
    blob = bytearray(open(os.path.join(base, doc["buffers"][0]["uri"]), "rb").read())
    for img in doc.get("images", []):
        data = open(os.path.join(base, img.pop("uri")), "rb").read()
        while len(blob) % 4:
            blob.append(0)
        doc["bufferViews"].append(dict(buffer=0, byteOffset=len(blob), byteLength=len(data)))
        img.update(bufferView=len(doc["bufferViews"]) - 1, mimeType="image/png")
        blob += data
    while len(blob) % 4:
        blob.append(0)
    doc["buffers"] = [dict(byteLength=len(blob))]
    js = json.dumps(doc).encode()
    js += b" " * (-len(js) % 4)
    import struct
    with open(glb_path, "wb") as f:
        f.write(struct.pack("<III", 0x46546C67, 2, 12 + 8 + len(js) + 8 + len(blob)))
        f.write(struct.pack("<II", len(js), 0x4E4F534A) + js)
        f.write(struct.pack("<II", len(blob), 0x004E4942) + bytes(blob))
    return glb_path


def _same_scene(a, b):
    for k in ("positions", "normals", "uvs", "indices"):
        assert a["mesh"][k].tobytes() == b["mesh"][k].tobytes(), k
    for k in ("primitives", "instances", "materials"):
        assert a[k].tobytes() == b[k].tobytes(), k
    assert len(a["textures"]) == len(b["textures"])
    for ta, tb in zip(a["textures"], b["textures"]):
        assert ta["srgb"] == tb["srgb"] and all(x.tobytes() == y.tobytes() for x, y in zip(ta["levels"], tb["levels"]))


def test_binary_container_and_embedded_images(tmp_path):
    """.glb (JSON chunk + BIN chunk, images behind buffer views) ingests to the same arrays as the .gltf it was packed from."""
    path, _ = _asset(tmp_path)
    a = gltf_ingest.finish(gltf_ingest.load_gltf(path))
    b = gltf_ingest.finish(gltf_ingest.load_gltf(_to_glb(path, str(tmp_path / "asset.glb"))))
    _same_scene(a, b)
    bad = tmp_path / "bad.glb"
    bad.write_bytes(b"nope" + bytes(40))
    with pytest.raises(ValueError):
        gltf_ingest.load_gltf(str(bad))


def test_interleaved_views_short_indices_and_data_uris(tmp_path):
    """Strided (interleaved) vertex views, 16-bit indices, normalised 16-bit uvs and base64 buffers — the accessor forms
    exporters commonly write — decode to the same floats as the plain layout."""
    import base64
    quad = scenes.quad_mesh(1.0)
    pos, nrm, uv, idx = (np.asarray(a) for a in quad)
    n = len(pos)
    inter = np.zeros((n, 8), f32)
    inter[:, 0:3], inter[:, 3:6] = pos, nrm
    uv16 = np.round(np.clip(uv, 0, 1) * 65535).astype(np.uint16)
    blob = inter.tobytes() + uv16.tobytes() + idx.astype(np.uint16).tobytes()
    o_uv, o_idx = inter.nbytes, inter.nbytes + uv16.nbytes
    doc = dict(asset=dict(version="2.0"), buffers=[dict(byteLength=len(blob), uri="data:application/octet-stream;base64," + base64.b64encode(blob).decode())],
               bufferViews=[dict(buffer=0, byteOffset=0, byteLength=inter.nbytes, byteStride=32),
                            dict(buffer=0, byteOffset=o_uv, byteLength=uv16.nbytes),
                            dict(buffer=0, byteOffset=o_idx, byteLength=2 * len(idx))],
               accessors=[dict(bufferView=0, byteOffset=0, componentType=5126, count=n, type="VEC3", min=pos.min(0).tolist(), max=pos.max(0).tolist()),
                          dict(bufferView=0, byteOffset=12, componentType=5126, count=n, type="VEC3"),
                          dict(bufferView=1, componentType=5123, normalized=True, count=n, type="VEC2"),
                          dict(bufferView=2, componentType=5123, count=len(idx), type="SCALAR")],
               meshes=[dict(primitives=[dict(attributes=dict(POSITION=0, NORMAL=1, TEXCOORD_0=2), indices=3)])],
               nodes=[dict(mesh=0)], scenes=[dict(nodes=[0])], scene=0)
    p = tmp_path / "interleaved.gltf"
    p.write_text(json.dumps(doc))
    m = gltf_ingest.finish(gltf_ingest.load_gltf(str(p)))
    assert m["mesh"]["positions"].tobytes() == pos.astype(f32).tobytes() and m["mesh"]["normals"].tobytes() == nrm.astype(f32).tobytes()
    assert m["mesh"]["indices"].dtype == np.uint32 and np.array_equal(m["mesh"]["indices"], idx)
    np.testing.assert_allclose(m["mesh"]["uvs"], np.clip(uv, 0, 1), atol=1.0 / 65535)
    assert len(m["materials"]) == 1 and m["materials"]["index_of_refraction"][0] == 1.5        # the default material
    # a sparse accessor (glTF 3.6.2.3; the gltf crate's iterators resolve it): two positions replaced on top of the base view
    repl = np.array([[5.0, 6.0, 7.0], [-1.0, -2.0, -3.0]], f32)
    extra = np.array([1, 3], np.uint16).tobytes() + repl.tobytes()
    blob2 = blob + extra
    doc2 = json.loads(json.dumps(doc))
    doc2["buffers"][0] = dict(byteLength=len(blob2), uri="data:application/octet-stream;base64," + base64.b64encode(blob2).decode())
    doc2["bufferViews"] += [dict(buffer=0, byteOffset=len(blob), byteLength=4), dict(buffer=0, byteOffset=len(blob) + 4, byteLength=24)]
    doc2["accessors"][0]["sparse"] = dict(count=2, indices=dict(bufferView=3, componentType=5123), values=dict(bufferView=4))
    doc2["accessors"][0]["min"], doc2["accessors"][0]["max"] = [-1.0, -2.0, -3.0], [5.0, 6.0, 7.0]
    p.write_text(json.dumps(doc2))
    sp = gltf_ingest.finish(gltf_ingest.load_gltf(str(p)))
    want = pos.astype(f32).copy()
    want[[1, 3]] = repl
    assert sp["mesh"]["positions"].tobytes() == want.tobytes()
    # ... and one without a base view starts from zeros
    del doc2["accessors"][1]["bufferView"]
    doc2["accessors"][1].pop("byteOffset", None)
    doc2["accessors"][1]["sparse"] = dict(count=2, indices=dict(bufferView=3, componentType=5123), values=dict(bufferView=4))
    p.write_text(json.dumps(doc2))
    sp = gltf_ingest.finish(gltf_ingest.load_gltf(str(p)))
    want_n = np.zeros_like(nrm, f32)
    want_n[[1, 3]] = repl
    assert sp["mesh"]["normals"].tobytes() == want_n.tobytes()
    # sparse indices out of range are refused, not misread
    blob3 = blob2 + np.array([3, 1], np.uint16).tobytes()                                    # not increasing
    doc2["buffers"][0] = dict(byteLength=len(blob3), uri="data:application/octet-stream;base64," + base64.b64encode(blob3).decode())
    doc2["bufferViews"].append(dict(buffer=0, byteOffset=len(blob2), byteLength=4))
    doc2["accessors"][0]["sparse"]["indices"] = dict(bufferView=5, componentType=5123)
    p.write_text(json.dumps(doc2))
    with pytest.raises(ValueError):
        gltf_ingest.load_gltf(str(p))
    # a primitive without indices is what the reference's loader panics on (model_loading.rs:100)
    doc3 = json.loads(json.dumps(doc))
    del doc3["meshes"][0]["primitives"][0]["indices"]
    p.write_text(json.dumps(doc3))
    with pytest.raises(ValueError):
        gltf_ingest.load_gltf(str(p))
    # and so is a node with non-uniform scale (model_loading.rs:449-458)
    doc["nodes"][0]["scale"] = [1.0, 2.0, 1.0]
    p.write_text(json.dumps(doc))
    with pytest.raises(ValueError):
        gltf_ingest.load_gltf(str(p))
