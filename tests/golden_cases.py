"""The golden cases of tests/golden/ (rendered by the unmodified reference) and how to
rebuild each scene + render parameters on our side."""
import ast
import os
import re

import numpy as np

from mitsuba2_b200 import abi, scenes

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def _rgb(xml, name, default=None):
    m = re.search(r'<rgb name="%s" value="([^"]+)"' % name, xml)
    return tuple(float(x) for x in m.group(1).split(",")) if m else default


def _float(xml, name, default):
    m = re.search(r'<float name="%s" value="([^"]+)"' % name, xml)
    return float(m.group(1)) if m else default


def _material(xml):
    """The XML fragments of tests/golden/make_golden.py -> SceneBuffers.add_* calls (plugin defaults as in the reference)."""
    if 'type="twosided"' in xml:
        inner = re.findall(r'<bsdf type="(?!twosided)[^"]+".*?</bsdf>', xml)

        def make(sb):
            front = _material(inner[0])(sb)
            back = _material(inner[1])(sb) if len(inner) > 1 else None
            return sb.make_twosided(front, back)
        return make
    if 'type="diffuse"' in xml:
        return lambda sb: sb.add_diffuse(_rgb(xml, "reflectance", (0.5, 0.5, 0.5)))
    if 'type="dielectric"' in xml:
        return lambda sb: sb.add_dielectric(_float(xml, "int_ior", 1.5046), _float(xml, "ext_ior", 1.000277))
    if 'type="roughdielectric"' in xml:
        return lambda sb: sb.add_roughdielectric(alpha=_float(xml, "alpha", 0.1), distribution="ggx" if 'value="ggx"' in xml else "beckmann",
                                                 int_ior=_float(xml, "int_ior", 1.5046), ext_ior=_float(xml, "ext_ior", 1.000277),
                                                 specular_transmittance=_rgb(xml, "specular_transmittance", (1, 1, 1)))
    if 'type="roughconductor"' in xml:
        return lambda sb: sb.add_roughconductor(alpha=_float(xml, "alpha", 0.1), distribution="ggx" if 'value="ggx"' in xml else "beckmann")
    if 'type="conductor"' in xml:
        return lambda sb: sb.add_conductor(eta=_rgb(xml, "eta", (0, 0, 0)), k=_rgb(xml, "k", (1, 1, 1)),
                                           specular_reflectance=_rgb(xml, "specular_reflectance", (1, 1, 1)))
    if 'type="plastic"' in xml:
        return lambda sb: sb.add_plastic(_rgb(xml, "diffuse_reflectance", (0.5, 0.5, 0.5)), _float(xml, "int_ior", 1.49),
                                         _float(xml, "ext_ior", 1.000277), _rgb(xml, "specular_reflectance"),
                                         nonlinear='name="nonlinear" value="true"' in xml)
    raise ValueError(xml)


def _tex_materials(z):
    """TEX_SPEC of tests/golden/make_golden.py -> textured BSDFs (to_uv = translate * scale, exact in fp32)."""
    spec = ast.literal_eval(str(z["tex_spec"]))
    out = {}
    for bsdf_id, (plugin, kind, kw) in spec.items():
        def make(sb, plugin=plugin, kind=kind, kw=kw):
            (sx, sy), (tx, ty) = kw["scale"], kw["translate"]
            to_uv = [[sx, 0, tx], [0, sy, ty], [0, 0, 1]]
            if kind == "checkerboard":
                t = sb.add_checkerboard(kw["color0"], kw["color1"], to_uv)
            else:
                t = sb.add_bitmap(z["tex_rgb"] if kind == "bitmap_rgb" else z["tex_gray"], kw["filter_type"], kw["wrap_mode"], to_uv)
            return sb.add_diffuse(texture=t) if plugin == "diffuse" else sb.add_plastic(int_ior=1.6, texture=t)
        out[bsdf_id] = make
    return out


def _materials(spec):
    return {name: _material(xml) for name, xml in spec.items()} if spec else None


def nopatch_meshes():
    """cbox meshes without the coplanar ceiling patch quad (faces 2,3 of mesh 2)."""
    z = dict(np.load(os.path.join(GOLDEN, "cbox_meshes.npz")))
    z["m2_faces"] = z["m2_faces"][:2]
    path = os.path.join(GOLDEN, "_cbox_meshes_nopatch.npz")
    np.savez(path, **z)
    return path


def load(name, host):
    """-> (SceneBuffers, B2RenderParams, reference rgb image)"""
    z = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))
    kw = ast.literal_eval(str(z["params"]))
    if kw.get("scene") == "veach":       # BASELINE configs[3] in miniature (mitsuba2_b200.scenes.veach_mis)
        sb = scenes.veach_mis(host, kw["width"], kw["height"])
        return sb, abi.render_params(max_depth=6, rr_depth=5, spp=kw["spp"], block_size=32), z["rgb"]
    env = None; drop = (0, 6) if kw.get("analytic") else ()      # analytic: luminaire + small box replaced by shapes
    if kw.get("point_light") in ("only", "spot"):
        drop = (0,)
    const_env = None
    if str(kw.get("envmap", "")).startswith("constant"):
        const_env = tuple(float(x) for x in z["const_radiance"]); drop = tuple(int(i) for i in z["env_dropped"])
    elif kw.get("envmap"):
        env = dict(rgb=z["env_rgb"], scale=float(z["env_scale"]), to_world=z["env_to_world"].reshape(4, 4))
        drop = tuple(int(i) for i in z["env_dropped"])
    sb = scenes.cbox(host, kw["width"], kw["height"], rfilter=kw.get("rfilter", "box"), envmap=env, drop_meshes=drop,
                     materials=_tex_materials(z) if kw.get("textures") else _materials(kw.get("materials")),
                     texcoords={i: z[f"uv_m{i}"] for i in range(8)} if kw.get("textures") else None,
                     meshes_npz=nopatch_meshes() if kw.get("no_patch") else None, lift=kw.get("lift", 0.0),
                     instances=kw.get("instances"), crop=kw.get("crop"))
    pl = ast.literal_eval(str(z["point_light"])) if kw.get("point_light") else None
    if kw.get("point_light") == "spot":
        sb.add_spot_light(pl["to_world"], pl["intensity"], pl["cutoff_angle"], pl["beam_width"])
    elif kw.get("point_light") == "directional":
        sb.add_directional_light(pl["to_world"], pl["irradiance"])      # before the constant emitter, like the XML
    elif kw.get("point_light"):
        sb.add_point_light(pl["position"], pl["intensity"])
    if const_env:                        # after the shapes, like the XML
        sb.add_constant_emitter(const_env)
    if kw.get("analytic"):              # SHAPES_SPEC of make_golden.py; emitters in XML order: rectangle, then sphere
        sp = ast.literal_eval(str(z["shapes_spec"]))
        ids = sb.bsdf_ids; light = ids["light"]
        r = sb.add_rectangle(sp["rect_light"]["to_world"], bsdf=light)
        sb.add_shape_emitter(r, sp["rect_light"]["radiance"])
        textured = kw["analytic"] == "textured"
        sphere_b = ids[sp["sphere"]["bsdf"]]
        if textured:        # checkerboard on the sphere, two-sided bitmap-textured rectangle (make_golden.shapes_xml)
            sphere_b = sb.add_diffuse(texture=sb.add_checkerboard((0.9, 0.8, 0.2), (0.1, 0.2, 0.6), [[8, 0, 0], [0, 4, 0], [0, 0, 1]]))
        sb.add_sphere(sp["sphere"]["center"], sp["sphere"]["radius"], bsdf=sphere_b)
        s2 = sb.add_sphere(sp["sphere_light"]["center"], sp["sphere_light"]["radius"], bsdf=light)
        sb.add_shape_emitter(s2, sp["sphere_light"]["radius"] and sp["sphere_light"]["radiance"])
        rect_b = ids[sp["rect_flip"]["bsdf"]]
        if textured:
            rect_b = sb.make_twosided(sb.add_diffuse(texture=sb.add_bitmap(z["tex_rgb"], "bilinear", "repeat", [[2, 0, 0], [0, 2, 0], [0, 0, 1]])))
        sb.add_rectangle(sp["rect_flip"]["to_world"], bsdf=rect_b, flip_normals=True)
    if kw.get("extra_emitter"):
        mi = sb.add_mesh(z["extra_positions"], z["extra_faces"], z["extra_normals"], bsdf=1)     # 'white'
        sb.add_area_emitter(mi, tuple(float(x) for x in z["extra_radiance"]))
    rp = abi.render_params(max_depth=kw.get("max_depth", 6), rr_depth=kw.get("rr_depth", 5), spp=kw["spp"],
                           samples_per_pass=kw.get("samples_per_pass") or 0, block_size=kw.get("block_size", 32),
                           seed=kw.get("seed", 0), sampler=kw.get("sampler", "independent"), jitter=kw.get("jitter", True))
    if kw.get("direct"):                 # DirectIntegrator with explicit sample counts (direct.cpp:78-100)
        rp.direct_emitter_samples, rp.direct_bsdf_samples = kw["direct"]
    return sb, rp, z["rgb"]


def names():
    return sorted(f[4:-4] for f in os.listdir(GOLDEN) if f.startswith("ref_") and f.endswith(".npz"))


def parity_stats(rgb, ref):
    """rel-L2 plus flip-robust statistics (DESIGN.md 'Float parity')."""
    rgb = np.asarray(rgb, np.float64); ref = np.asarray(ref, np.float64)
    d = np.abs(rgb - ref).max(axis=2)
    rel = np.abs(rgb - ref) / np.maximum(np.abs(ref), 1e-3 * ref.mean())
    flipped = d > 1e-3 * ref.mean()
    clean = ~flipped
    return dict(rel_l2=float(np.linalg.norm(rgb - ref) / np.linalg.norm(ref)),
                rel_l2_unflipped=float(np.linalg.norm((rgb - ref)[clean]) / np.linalg.norm(ref[clean])),
                median_rel=float(np.median(rel)), p99_rel=float(np.percentile(rel, 99)),
                flipped_pixels=int(flipped.sum()), pixels=int(d.size), max_abs=float(d.max()))
