"""The frame kernels' light loop, element by element (tr_eval_point_light), against the oracle.

tr_eval_basic_brdf / tr_eval_transmission_btdf test the all-exact contract wrappers; the kernels that make the benchmark
number run the fast regime with adaptive exactness (tr_device_pbr.cuh: light_lean).  This file drives THAT code — the same
device functions, through the C ABI — with one (pixel, point light) pair per element and checks every element, not an
aggregate: glam-pbr/src/lib.rs:12-23 (light_direction_and_attenuation), :377-423 (basic_brdf), :200-233
(transmission_btdf); loop bodies shader/src/lighting.rs:58-92, 179-216.

Populations: uniformly random directions, highlight peaks (l within a few mrad of the mirror direction: the exact-regime
patch), grazing lights (n.l ~ 0), rim pixels (n.v ~ 0) and back-facing normals (n.v < 0, clamped to EPSILON by Dot::new).
"""
import numpy as np
import pytest

from pipeline import rel_l2
from transmission_renderer_b200 import Renderer, abi

pytestmark = pytest.mark.gpu
f32 = np.float32


def _unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def _materials(rng, n, rough_lo):
    m = np.zeros(n, dtype=abi.material_params)
    m["diffuse_colour"] = rng.uniform(0.05, 1.0, (n, 3))
    m["metallic"] = rng.choice([0.0, 1.0, 0.3], n)
    m["perceptual_roughness"] = rng.uniform(rough_lo, 1.0, n)
    m["index_of_refraction"] = rng.uniform(1.0, 2.2, n)
    m["specular_colour"] = rng.uniform(0.2, 1.0, (n, 3))
    m["specular_factor"] = rng.uniform(0.0, 1.0, n)
    return m


def _case(rng, n, kind, rough_lo=0.05):
    nrm = _unit(rng.standard_normal((n, 3)))
    t = _unit(np.cross(nrm, rng.standard_normal((n, 3))))           # a tangent
    b = np.cross(nrm, t)
    phi = rng.uniform(0, 2 * np.pi, (n, 1))
    cos_v = rng.uniform(0.02, 1.0, (n, 1))
    if kind == "rim":
        cos_v = rng.uniform(-2e-3, 2e-3, (n, 1))
    elif kind == "backfacing":
        cos_v = rng.uniform(-1.0, -0.01, (n, 1))
    sin_v = np.sqrt(np.maximum(0.0, 1 - cos_v ** 2))
    view = cos_v * nrm + sin_v * (np.cos(phi) * t + np.sin(phi) * b)
    if kind == "highlight":      # mirror direction + a few milliradians
        ldir = _unit(2 * np.sum(nrm * view, axis=1, keepdims=True) * nrm - view + rng.standard_normal((n, 3)) * 10.0 ** rng.uniform(-4, -1.5, (n, 1)))
    elif kind == "highlight_t":  # the transmission lobe's peak: the mirrored light l' = l - 2 (n.l) n is the mirror direction
        m = _unit(2 * np.sum(nrm * view, axis=1, keepdims=True) * nrm - view + rng.standard_normal((n, 3)) * 10.0 ** rng.uniform(-4, -1.5, (n, 1)))
        ldir = m - 2 * np.sum(nrm * m, axis=1, keepdims=True) * nrm
    elif kind == "grazing":
        psi = rng.uniform(0, 2 * np.pi, (n, 1))
        ldir = _unit(rng.uniform(-3e-3, 3e-3, (n, 1)) * nrm + np.cos(psi) * t + np.sin(psi) * b)
    else:
        ldir = _unit(rng.standard_normal((n, 3)))
    dist = 10.0 ** rng.uniform(-1.3, 1.3, (n, 1))                    # 5 cm ... 20 m
    pos = rng.uniform(-30, 30, (n, 3))
    p = np.zeros(n, dtype=abi.point_light_params)
    p["normal"] = _unit(nrm.astype(f32))
    p["view"] = _unit(view.astype(f32))
    p["position"] = pos
    p["light_position"] = pos + ldir * dist
    p["light_colour"] = rng.uniform(0.5, 50.0, (n, 3))
    p["material_params"] = _materials(rng, n, rough_lo)
    return p


def _conditioning(p):
    """float64 view of the inputs: n.l, n.v, and the squared lengths of the two halfway vectors / 2 (1 + v.l, 1 + v.l')."""
    n, v = p["normal"].astype(np.float64), p["view"].astype(np.float64)
    l = _unit(p["light_position"].astype(np.float64) - p["position"].astype(np.float64))
    nol, nov, vol = np.sum(n * l, 1), np.sum(n * v, 1), np.sum(v * l, 1)
    return nol, nov, 1.0 + vol, 1.0 + vol - 2.0 * nol * nov


def _ulp_sensitivity(oracle, p, ref):
    """max |reference(inputs nudged by one ulp) - reference(inputs)| per element and lobe: how far the REFERENCE's own result
    moves when one input component moves by one unit in the last place.  Nudged: each light-position component, one
    component of the normal and of the view vector (six re-evaluations)."""
    dev = {k: np.zeros(ref[k].shape, np.float64) for k in ("diffuse", "specular", "transmission")}
    for field, comp in (("light_position", 0), ("light_position", 1), ("light_position", 2), ("normal", 0), ("view", 1), ("position", 2)):
        q = p.copy()
        x = q[field][:, comp]
        q[field][:, comp] = np.nextafter(x, np.where(x > 0, np.float32(np.inf), np.float32(-np.inf)).astype(np.float32))
        r = oracle.eval_point_light(q)
        for k in dev:
            d = np.abs(r[k].astype(np.float64) - ref[k].astype(np.float64))
            dev[k] = np.maximum(dev[k], np.where(np.isfinite(d), d, np.inf))
    return dev


# Per element and channel:  |got - ref| <= RTOL |ref| + ATOL x (light colour x attenuation + the lobe's largest channel) + 4 x S,
# where S is the reference's own sensitivity to a one-ulp nudge of its inputs (_ulp_sensitivity).  Where the lobes are regular S is ~1e-7 |ref|
# and the bound is the relative tolerance; where the reference's value is rounding noise — the halfway vector v + l (or
# v + l' of the mirrored light) vanishes and `Halfway::new` normalises a zero vector (glam-pbr lib.rs:64-68), or n.v / n.l are
# clamped to EPSILON and 1 / (n.l n.v) amplifies every ulp by 1e7 (`Dot::new`, lib.rs:92-99) — the bound widens by exactly
# as much as the reference itself is undetermined.  ATOL: the fp32 resolution of the unit-vector dot products (8 ulp of 1);
# the three channels of a lobe share D V, so a channel that (1 - F) cancels to a millionth of the others is held to
# their scale.  Not judged at all (only counted): elements whose halfway vector is shorter than 0.015 (1 + v.l < 1e-4 for the
# reflection lobe, 1 + v.l' < 1e-4 for the transmission lobe) — there the reference's n.h and v.h are quotients of two
# rounding residues, and nudging the inputs by an ulp does not reveal it.
RTOL, ATOL, DEGENERATE = 1e-4, 1e-6, 1e-4


@pytest.mark.parametrize("kind", ["random", "highlight", "highlight_t", "grazing", "rim", "backfacing"])
def test_point_light_per_element(oracle, kind):
    rng = np.random.default_rng({"random": 1, "highlight": 2, "highlight_t": 3, "grazing": 4, "rim": 5, "backfacing": 6}[kind])
    n = 250000 if kind == "random" else 100000
    p = _case(rng, n, kind)
    with Renderer(64, 64) as r:
        got = r.eval_point_light(p)
    ref = oracle.eval_point_light(p)
    sens = _ulp_sensitivity(oracle, p, ref)
    d2 = np.sum((p["light_position"].astype(np.float64) - p["position"]) ** 2, axis=1, keepdims=True)
    scale = p["light_colour"].astype(np.float64) / d2          # colour x attenuation
    report = []
    _, _, opv, opvt = _conditioning(p)
    judged = {"diffuse": np.ones(n, bool), "specular": opv >= DEGENERATE, "transmission": opvt >= DEGENERATE}
    for k in ("diffuse", "specular", "transmission"):
        g, f = got[k].astype(np.float64), ref[k].astype(np.float64)
        finite = np.isfinite(f).all(axis=1) & np.isfinite(g).all(axis=1) & np.isfinite(sens[k]).all(axis=1)
        assert finite.mean() > 0.99, (kind, k, finite.mean())
        finite &= judged[k]
        assert finite.sum() > 2000, (kind, k, int(finite.sum()))
        tol = RTOL * np.abs(f) + ATOL * (scale + np.abs(f).max(axis=1, keepdims=True)) + 4.0 * sens[k]
        excess = (np.abs(g - f) / tol)[finite]
        regular = (4.0 * sens[k] <= RTOL * np.abs(f) + ATOL * scale).all(axis=1)[finite]   # bound not widened by the sensitivity term
        bad = (excess > 1.0).any(axis=1)
        l2 = rel_l2(got[k][finite][regular], ref[k][finite][regular])
        report.append(f"{k}: worst {excess.max():.2f} x tolerance, {bad.sum()} of {len(bad)} judged elements over ({n - len(bad)} degenerate / non-finite); "
                      f"{regular.mean():.3f} regular (rel-L2 there {l2:.1e})")
        if bad.sum():   # show what the offending elements look like before failing
            idx = np.nonzero(finite)[0][np.argsort(-excess.max(axis=1))[:4]]
            nol, nov, opv, opvt = _conditioning(p[idx])
            for j, i in enumerate(idx):
                print(f"  worst[{k}] #{i}: got {got[k][i]} ref {ref[k][i]} sens {sens[k][i]} n.l {nol[j]:.3e} n.v {nov[j]:.3e} 1+v.l {opv[j]:.3e} "
                      f"1+v.l' {opvt[j]:.3e} rough {p['material_params']['perceptual_roughness'][i]:.3f} ior {p['material_params']['index_of_refraction'][i]:.3f} "
                      f"metallic {p['material_params']['metallic'][i]:.1f} dist {np.sqrt(d2[i, 0]):.3f}")
        assert bad.sum() == 0, (kind, k, float(excess.max()), int(bad.sum()))
        assert l2 < 1e-4, (kind, k, l2)   # the north-star tolerance, on the regular elements of the population
    print(f"point light [{kind}] " + " | ".join(report))
