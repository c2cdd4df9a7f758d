"""glTF 2.0 ingest into the flat arrays the renderer consumes (row N4 of SURVEY.md 8f).

Host-side mirror of the reference's loader, `load_gltf` in /root/reference/src/model_loading.rs:13-338 and `NodeTree`
(:438-484): node hierarchy flattened to one `Similarity` per mesh node, positions / normals / uvs / indices appended to
shared arrays, ONE `Instance` and ONE `PrimitiveInfo` per glTF primitive (:136-160), bounding sphere from the accessor's
bounding box (:148-155), draw buffer chosen from alpha mode x KHR_materials_transmission (:68-78), `MaterialInfo` from
pbrMetallicRoughness plus the KHR ior / transmission / volume / specular extensions with the loader's defaults
(:231-333), images bound once per (image, sRGB-ness) pair (:166-230).  The result feeds `Renderer.set_*` directly.

Not the hot path: this is data-format plumbing beside it, in plain Python/numpy (the reference does it on the CPU too).
glTF-Sample-Models is not available offline, so the tests write their own assets with `write_gltf`.
"""
import base64
import json
import os
import struct

import numpy as np

from . import abi, host
from .scenes import TEX_SLOTS, make_mips

f32 = np.float32

_COMPONENT = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_WIDTH = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}


# --------------------------------------------------------------------------- Similarity (shared-structs lib.rs:183-241)
def quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz], f32)


class Similarity:
    def __init__(self, translation=(0, 0, 0), rotation=(0, 0, 0, 1), scale=1.0):
        self.translation = np.asarray(translation, f32)
        self.rotation = np.asarray(rotation, f32)
        self.scale = f32(scale)

    def apply(self, v):  # Mul<Vec3>: translation + scale * (rotation * v)
        return (self.translation + self.scale * host.quat_rotate(self.rotation, np.asarray(v, f32))).astype(f32)

    def __mul__(self, child):  # Mul<Similarity>, lib.rs:222-232
        return Similarity(self.apply(child.translation), quat_mul(self.rotation, child.rotation), self.scale * child.scale)


def _decompose(m):
    """gltf `Transform::decomposed` of a column-major 4x4: translation, rotation quaternion (x, y, z, w), scale."""
    m = np.asarray(m, np.float64).reshape(4, 4).T  # math convention
    t = m[:3, 3]
    s = np.linalg.norm(m[:3, :3], axis=0)
    if np.linalg.det(m[:3, :3]) < 0:
        s = -s
    r = m[:3, :3] / s
    tr = r[0, 0] + r[1, 1] + r[2, 2]
    if tr > 0:
        k = np.sqrt(tr + 1.0) * 2
        q = [(r[2, 1] - r[1, 2]) / k, (r[0, 2] - r[2, 0]) / k, (r[1, 0] - r[0, 1]) / k, 0.25 * k]
    elif r[0, 0] > r[1, 1] and r[0, 0] > r[2, 2]:
        k = np.sqrt(1.0 + r[0, 0] - r[1, 1] - r[2, 2]) * 2
        q = [0.25 * k, (r[0, 1] + r[1, 0]) / k, (r[0, 2] + r[2, 0]) / k, (r[2, 1] - r[1, 2]) / k]
    elif r[1, 1] > r[2, 2]:
        k = np.sqrt(1.0 + r[1, 1] - r[0, 0] - r[2, 2]) * 2
        q = [(r[0, 1] + r[1, 0]) / k, 0.25 * k, (r[1, 2] + r[2, 1]) / k, (r[0, 2] - r[2, 0]) / k]
    else:
        k = np.sqrt(1.0 + r[2, 2] - r[0, 0] - r[1, 1]) * 2
        q = [(r[0, 2] + r[2, 0]) / k, (r[1, 2] + r[2, 1]) / k, 0.25 * k, (r[1, 0] - r[0, 1]) / k]
    return t, q, s


class NodeTree:
    """model_loading.rs:438-484: per node its local Similarity and its parent; `transform_of` walks to the root."""

    def __init__(self, nodes):
        self.local, self.parent = [], [-1] * len(nodes)
        for i, node in enumerate(nodes):
            if "matrix" in node:
                t, q, s = _decompose(node["matrix"])
            else:
                t, q, s = node.get("translation", (0, 0, 0)), node.get("rotation", (0, 0, 0, 1)), node.get("scale", (1, 1, 1))
            s = np.asarray(s, f32)
            eps = np.finfo(f32).eps * 10
            if abs(s[0] - s[1]) > eps * max(1.0, abs(s[0])) or abs(s[0] - s[2]) > eps * max(1.0, abs(s[0])):
                raise ValueError(f"node {i}: non-uniform scale {s} (the reference asserts uniform scale, model_loading.rs:449-458)")
            self.local.append(Similarity(t, q, s[0]))
            for child in node.get("children", []):
                self.parent[child] = i

    def transform_of(self, index):
        total = Similarity()
        while index != -1:
            total = self.local[index] * total
            index = self.parent[index]
        return total


# --------------------------------------------------------------------------- accessors
class _Buffers:
    def __init__(self, doc, base_dir, glb_chunk=None):
        self.doc, self.blobs = doc, []
        for b in doc.get("buffers", []):
            uri = b.get("uri")
            if uri is None:
                self.blobs.append(glb_chunk)
            elif uri.startswith("data:"):
                self.blobs.append(base64.b64decode(uri.split(",", 1)[1]))
            else:
                self.blobs.append(open(os.path.join(base_dir, uri), "rb").read())

    def view_bytes(self, view_index):
        v = self.doc["bufferViews"][view_index]
        blob = self.blobs[v["buffer"]]
        off = v.get("byteOffset", 0)
        return blob[off:off + v["byteLength"]], v.get("byteStride")

    def read(self, accessor_index):
        a = self.doc["accessors"][accessor_index]
        dt, width, count = np.dtype(_COMPONENT[a["componentType"]]), _WIDTH[a["type"]], a["count"]
        elem = dt.itemsize * width

        def dense(view_index, byte_offset, n, dtype, w):
            data, stride = self.view_bytes(view_index)
            e = dtype.itemsize * w
            if stride in (None, 0, e):
                return np.frombuffer(data, dtype=dtype, count=n * w, offset=byte_offset).reshape(n, w)
            raw = np.frombuffer(data, dtype=np.uint8)
            idx = byte_offset + np.arange(n)[:, None] * stride + np.arange(e)[None, :]
            return raw[idx].view(dtype).reshape(n, w)

        # the gltf crate's accessor iterators (what model_loading.rs reads through) resolve sparse accessors: zeros or the
        # base view, then `count` elements replaced at the given indices (glTF 2.0, 5.1.4 / 3.6.2.3)
        if "bufferView" in a:
            out = dense(a["bufferView"], a.get("byteOffset", 0), count, dt, width)
        else:
            out = np.zeros((count, width), dt)
        if "sparse" in a:
            sp = a["sparse"]
            n = sp["count"]
            si, sv = sp["indices"], sp["values"]
            where = dense(si["bufferView"], si.get("byteOffset", 0), n, np.dtype(_COMPONENT[si["componentType"]]), 1).reshape(-1).astype(np.int64)
            if n and (where.max() >= count or (np.diff(where) <= 0).any()):
                raise ValueError(f"accessor {accessor_index}: sparse indices must be strictly increasing and below {count}")
            vals = dense(sv["bufferView"], sv.get("byteOffset", 0), n, dt, width)
            out = np.array(out, copy=True)
            out[where] = vals
        if a.get("normalized") and dt.kind in "iu":
            out = out.astype(np.float32) / float(np.iinfo(dt).max)
            if dt.kind == "i":
                out = np.maximum(out, -1.0)
        return out


def _load_image(doc, bufs, base_dir, image_index):
    from PIL import Image
    import io
    img = doc["images"][image_index]
    if "uri" in img:
        uri = img["uri"]
        data = base64.b64decode(uri.split(",", 1)[1]) if uri.startswith("data:") else open(os.path.join(base_dir, uri), "rb").read()
    else:
        data, _ = bufs.view_bytes(img["bufferView"])
    return np.asarray(Image.open(io.BytesIO(data)).convert("RGBA"), np.uint8)  # R8G8B8 -> R8G8B8A8, model_loading.rs:36-52


# --------------------------------------------------------------------------- the loader
def load_gltf(path, base_transform=None, roughness_override=None, into=None):
    """Returns (or extends `into`) dict(mesh=dict(positions, normals, uvs, indices), primitives, instances, materials,
    textures, max_draw_counts) — ModelStagingBuffers + ImageManager + MaxDrawCounts of the reference (main.rs:2495-2560)."""
    base_transform = base_transform or Similarity()
    base_dir = os.path.dirname(os.path.abspath(path))
    glb_chunk = None
    if path.endswith(".glb"):
        raw = open(path, "rb").read()
        magic, _, _ = struct.unpack_from("<III", raw, 0)
        if magic != 0x46546C67:
            raise ValueError("not a GLB file")
        jl, _ = struct.unpack_from("<II", raw, 12)
        doc = json.loads(raw[20:20 + jl])
        if 20 + jl < len(raw):
            bl, _ = struct.unpack_from("<II", raw, 20 + jl)
            glb_chunk = raw[28 + jl:28 + jl + bl]
    else:
        doc = json.load(open(path))
    bufs = _Buffers(doc, base_dir, glb_chunk)
    m = into or dict(pos=[], nrm=[], uv=[], idx=[], prims=[], inst=[], mats=[], textures=[], n_vertices=0, n_indices=0,
                     max_draw_counts=dict(opaque=0, alpha_clip=0, transmission=0, transmission_alpha_clip=0))
    materials = doc.get("materials", [])
    if not materials:
        # `material.index().unwrap_or(0)` (model_loading.rs:96) points a material-less primitive at entry 0, which does not
        # exist in a file without materials (the reference would read past its buffer); bind the glTF default material there.
        materials = [{}]
    n_existing_materials = sum(len(x) for x in m["mats"])
    tree = NodeTree(doc.get("nodes", []))

    for node_index, node in enumerate(doc.get("nodes", [])):
        if "mesh" not in node:
            continue
        transform = base_transform * tree.transform_of(node_index)
        for prim in doc["meshes"][node["mesh"]]["primitives"]:
            mat = materials[prim["material"]] if "material" in prim else {}
            ext = mat.get("extensions", {})
            has_transmission = "KHR_materials_transmission" in ext
            mode = mat.get("alphaMode", "OPAQUE")
            bucket = {("OPAQUE", False): 0, ("MASK", False): 1, ("OPAQUE", True): 2, ("MASK", True): 3}.get((mode, has_transmission), 0)
            m["max_draw_counts"][("opaque", "alpha_clip", "transmission", "transmission_alpha_clip")[bucket]] += 1
            tinfo = mat.get("pbrMetallicRoughness", {}).get("baseColorTexture", {})
            uv_scale = np.asarray(tinfo.get("extensions", {}).get("KHR_texture_transform", {}).get("scale", (1, 1)), f32)
            material_id = prim.get("material", 0) + n_existing_materials
            attrs = prim["attributes"]
            if "indices" not in prim:   # `reader.read_indices().unwrap()`, model_loading.rs:100: the reference panics here
                raise ValueError(f"node {node_index}: primitive without indices (the reference's loader requires indexed primitives)")
            if prim.get("mode", 4) != 4:
                raise ValueError(f"node {node_index}: primitive mode {prim.get('mode')} (only triangle lists are drawn)")
            idx = bufs.read(prim["indices"]).reshape(-1).astype(np.uint32)
            pos = bufs.read(attrs["POSITION"]).astype(f32)
            nrm = bufs.read(attrs["NORMAL"]).astype(f32)
            uv = (bufs.read(attrs["TEXCOORD_0"]).astype(f32) * uv_scale).astype(f32) if "TEXCOORD_0" in attrs \
                else np.zeros((len(pos), 2), f32)                                   # model_loading.rs:121-134
            acc = doc["accessors"][attrs["POSITION"]]
            lo, hi = np.asarray(acc["min"], f32), np.asarray(acc["max"], f32)       # primitive.bounding_box()
            p = np.zeros(1, dtype=abi.primitive_info)
            p["packed_bounding_sphere"][0] = (*((lo + hi) / f32(2)), f32(np.linalg.norm((hi - lo).astype(f32)) / f32(2)))
            p["draw_buffer_index"] = bucket
            p["index_count"] = len(idx)
            p["first_index"] = m["n_indices"]
            p["first_instance"] = len(m["inst"])
            inst = np.zeros(1, dtype=abi.instance)
            inst["translation_and_scale"][0] = (*transform.translation, transform.scale)
            inst["rotation"][0] = transform.rotation
            inst["primitive_id"] = len(m["prims"])
            inst["material_id"] = material_id
            m["idx"].append(idx + np.uint32(m["n_vertices"]))
            m["pos"].append(pos)
            m["nrm"].append(nrm)
            m["uv"].append(uv)
            m["prims"].append(p)
            m["inst"].append(inst)
            m["n_vertices"] += len(pos)
            m["n_indices"] += len(idx)

    # materials, model_loading.rs:169-333
    image_to_id = {}

    def texture(info, requirement):
        if info is None:
            return -1
        image_index = doc["textures"][info["index"]]["source"]
        if requirement == "dont_care":                      # read from alpha: reuse an sRGB binding if there is one
            if (image_index, True) in image_to_id:
                return image_to_id[(image_index, True)]
            srgb = False
        else:
            srgb = requirement == "srgb"
        key = (image_index, srgb)
        if key not in image_to_id:
            rgba = _load_image(doc, bufs, base_dir, image_index)
            image_to_id[key] = len(m["textures"])
            m["textures"].append(dict(levels=make_mips(rgba, srgb), srgb=srgb))
        return image_to_id[key]

    out = abi.default_material(len(materials))
    for i, mat in enumerate(materials):
        pbr = mat.get("pbrMetallicRoughness", {})
        ext = mat.get("extensions", {})
        tr_, vol, spec = ext.get("KHR_materials_transmission"), ext.get("KHR_materials_volume"), ext.get("KHR_materials_specular")
        t = out["textures"][i]
        t[TEX_SLOTS["diffuse"]] = texture(pbr.get("baseColorTexture"), "srgb")
        t[TEX_SLOTS["metallic_roughness"]] = texture(pbr.get("metallicRoughnessTexture"), "linear")
        t[TEX_SLOTS["normal_map"]] = texture(mat.get("normalTexture"), "linear")
        t[TEX_SLOTS["emissive"]] = texture(mat.get("emissiveTexture"), "srgb")
        t[TEX_SLOTS["occlusion"]] = texture(mat.get("occlusionTexture"), "linear")
        t[TEX_SLOTS["transmission"]] = texture(tr_.get("transmissionTexture") if tr_ else None, "linear")
        t[TEX_SLOTS["thickness"]] = texture(vol.get("thicknessTexture") if vol else None, "linear")
        t[TEX_SLOTS["specular_colour"]] = texture(spec.get("specularColorTexture") if spec else None, "srgb")
        t[TEX_SLOTS["specular"]] = texture(spec.get("specularTexture") if spec else None, "dont_care")
        out["metallic_factor"][i] = pbr.get("metallicFactor", 1.0)
        out["roughness_factor"][i] = roughness_override if roughness_override is not None else pbr.get("roughnessFactor", 1.0)
        out["alpha_clipping_cutoff"][i] = mat.get("alphaCutoff", 0.5)
        out["diffuse_factor"][i] = pbr.get("baseColorFactor", (1, 1, 1, 1))
        out["emissive_factor"][i, :3] = mat.get("emissiveFactor", (0, 0, 0))
        out["normal_map_scale"][i] = mat["normalTexture"].get("scale", 1.0) if "normalTexture" in mat else 0.0   # unwrap_or_default
        out["occlusion_strength"][i] = mat["occlusionTexture"].get("strength", 1.0) if "occlusionTexture" in mat else 1.0
        out["index_of_refraction"][i] = ext.get("KHR_materials_ior", {}).get("ior", 1.5)
        out["transmission_factor"][i] = tr_.get("transmissionFactor", 0.0) if tr_ is not None else 0.0
        out["thickness_factor"][i] = vol.get("thicknessFactor", 0.0) if vol is not None else 0.0
        out["attenuation_distance"][i] = f32(vol.get("attenuationDistance", np.inf)) * base_transform.scale if vol is not None else np.inf
        out["attenuation_colour"][i, :3] = vol.get("attenuationColor", (1, 1, 1)) if vol is not None else (1, 1, 1)
        out["specular_factor"][i] = spec.get("specularFactor", 1.0) if spec is not None else 1.0
        out["specular_colour_factor"][i, :3] = spec.get("specularColorFactor", (1, 1, 1)) if spec is not None else (1, 1, 1)
    m["mats"].append(out)
    return m


def finish(m):
    """Concatenate the staging lists into the arrays `Renderer.set_mesh / set_primitives / set_instances / set_materials`
    take (ModelStagingBuffers::upload, src/main.rs:2516-2557)."""
    return dict(mesh=dict(positions=np.concatenate(m["pos"]), normals=np.concatenate(m["nrm"]), uvs=np.concatenate(m["uv"]),
                          indices=np.concatenate(m["idx"])),
                primitives=np.concatenate(m["prims"]), instances=np.concatenate(m["inst"]),
                materials=np.concatenate(m["mats"]) if m["mats"] else abi.default_material(0), textures=m["textures"],
                max_draw_counts=m["max_draw_counts"])


# --------------------------------------------------------------------------- a writer, so the tests can make assets offline
def write_gltf(path, nodes, meshes, materials, images=()):
    """nodes: list of dict(mesh=?, translation=?, rotation=?, scale=?, children=?); meshes: list of lists of
    dict(positions, normals, uvs or None, indices, material); materials: glTF material dicts (texture infos refer to
    textures[i] == images[i]); images: list of (h, w, 4) uint8 arrays, written as PNG files next to the .gltf."""
    from PIL import Image
    base_dir = os.path.dirname(os.path.abspath(path))
    stem = os.path.splitext(os.path.basename(path))[0]
    blob, views, accessors = bytearray(), [], []

    def add(arr, target, kind, comp, with_bounds=False):
        arr = np.ascontiguousarray(arr)
        while len(blob) % 4:
            blob.append(0)
        views.append(dict(buffer=0, byteOffset=len(blob), byteLength=arr.nbytes, target=target))
        blob.extend(arr.tobytes())
        acc = dict(bufferView=len(views) - 1, componentType=comp, count=len(arr), type=kind)
        if with_bounds:
            acc["min"], acc["max"] = [float(x) for x in arr.min(axis=0)], [float(x) for x in arr.max(axis=0)]
        accessors.append(acc)
        return len(accessors) - 1

    gl_meshes = []
    for prims in meshes:
        out = []
        for p in prims:
            attrs = dict(POSITION=add(np.asarray(p["positions"], f32), 34962, "VEC3", 5126, True),
                         NORMAL=add(np.asarray(p["normals"], f32), 34962, "VEC3", 5126))
            if p.get("uvs") is not None:
                attrs["TEXCOORD_0"] = add(np.asarray(p["uvs"], f32), 34962, "VEC2", 5126)
            d = dict(attributes=attrs, indices=add(np.asarray(p["indices"], np.uint32), 34963, "SCALAR", 5125))
            if p.get("material") is not None:
                d["material"] = p["material"]
            out.append(d)
        gl_meshes.append(dict(primitives=out))
    image_entries = []
    for i, img in enumerate(images):
        name = f"{stem}_image{i}.png"
        Image.fromarray(np.asarray(img, np.uint8), "RGBA").save(os.path.join(base_dir, name))
        image_entries.append(dict(uri=name))
    with open(os.path.join(base_dir, stem + ".bin"), "wb") as f:
        f.write(bytes(blob))
    used = sorted({k for mat in materials for k in mat.get("extensions", {})})
    doc = dict(asset=dict(version="2.0"), scene=0, scenes=[dict(nodes=[i for i in range(len(nodes)) if not any(
        i in n.get("children", []) for n in nodes)])], nodes=nodes, meshes=gl_meshes, materials=materials,
        buffers=[dict(uri=stem + ".bin", byteLength=len(blob))], bufferViews=views, accessors=accessors)
    if images:
        doc["images"] = image_entries
        doc["textures"] = [dict(source=i) for i in range(len(images))]
    if used:
        doc["extensionsUsed"] = used
    with open(path, "w") as f:
        json.dump(doc, f)
    return path
