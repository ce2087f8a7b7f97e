#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference for the cbox OBJ data and
oracle/_ref/dist/mitsuba built by oracle/build_ref.sh). The outputs are small, committed,
and are what pins the CPU oracle (oracle/b2o.c) to the reference:

  cbox_meshes.npz          the 8 cbox meshes exactly as the reference's OBJ loader produces
                           them (src/shapes/obj.cpp:150-346: strtof, (v,vt,vn)-keyed vertex
                           dedup, fan triangulation, to_world applied at load; vertex normals
                           from the file or angle-weighted, src/librender/mesh.cpp:200-256)
  ref_<name>.npz           fp32 linear-RGB images rendered by `mitsuba -m scalar_rgb` with
                           pinned block_size and component_format=float32 / file_format=pfm,
                           plus the exact CLI parameters used.

Usage:  python tests/golden/make_golden.py [--only NAME]
"""
import argparse
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("B2_REFERENCE", "/root/reference")
DIST = os.path.join(ROOT, "oracle", "_ref", "dist")
CBOX = os.path.join(REF, "resources", "data", "scenes", "cbox")

f32 = np.float32


def load_obj_like_reference(path, translate=(0.0, 0.0, 0.0)):
    """Mirror of OBJMesh (src/shapes/obj.cpp:150-346) for v / vt / vn / f records."""
    verts, norms, uvs = [], [], []
    vmap = {}          # key (v, vt, vn) -> id, ids handed out in order of first use
    keys = []
    tris = []
    t = np.array(translate, dtype=f32)
    for line in open(path):
        cur = line.strip()
        if cur.startswith("v "):
            p = np.array([f32(x) for x in cur.split()[1:4]], dtype=f32)
            # Transform4f::translate().transform_affine(p): fmadd(col_i, p_i, col3) == p + t in fp32
            verts.append((p + t).astype(f32))
        elif cur.startswith("vn "):
            n = np.array([f32(x) for x in cur.split()[1:4]], dtype=f32)
            n = n * f32(1.0) / np.sqrt((n * n).sum(dtype=f32), dtype=f32)
            norms.append(n.astype(f32))
        elif cur.startswith("vt "):
            uv = np.array([f32(x) for x in cur.split()[1:3]], dtype=f32)
            uv[1] = f32(1.0) - uv[1]                       # flip_tex_coords (default true), obj.cpp:199-200
            uvs.append(uv)
        elif cur.startswith("f "):
            ids = []
            for tok in cur.split()[1:]:
                parts = tok.split("/")
                key = [0, 0, 0]
                for i, s in enumerate(parts[:3]):
                    key[i] = int(s) if s else 0
                key = tuple(key)
                if key not in vmap:
                    vmap[key] = len(keys); keys.append(key)
                ids.append(vmap[key])
            tri = [0, 0, 0]
            for vi, idx in enumerate(ids):
                if vi < 3:
                    tri[vi] = idx
                else:
                    tri[1] = tri[2]; tri[2] = idx
                if vi >= 2:
                    tris.append(tuple(tri))
    positions = np.array([verts[k[0] - 1] for k in keys], dtype=f32)
    faces = np.array(tris, dtype=np.uint32)
    if norms:
        normals = np.array([norms[k[2] - 1] for k in keys], dtype=f32)
    else:
        normals = recompute_vertex_normals(positions, faces)
    if uvs:
        return positions, normals, faces, np.array([uvs[k[1] - 1] for k in keys], dtype=f32)
    return positions, normals, faces


def _cross(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], dtype=f32)


def _normalize(v):
    return (v * (f32(1.0) / np.sqrt(np.dot(v, v).astype(f32), dtype=f32))).astype(f32)


def _unit_angle(u, v):
    # include/mitsuba/core/vector.h unit_angle(): numerically robust angle between unit vectors
    d = f32(np.dot(u, v))
    w = v - (u if d >= 0 else -u)
    temp = f32(2.0) * np.arcsin(f32(0.5) * np.sqrt(np.dot(w, w).astype(f32), dtype=f32), dtype=f32)
    return temp if d >= 0 else f32(np.pi) - temp


def recompute_vertex_normals(positions, faces):
    """Mesh::recompute_vertex_normals, scalar branch (src/librender/mesh.cpp:200-256)."""
    normals = np.zeros_like(positions)
    for f in faces:
        v = [positions[f[0]], positions[f[1]], positions[f[2]]]
        s0, s1 = v[1] - v[0], v[2] - v[0]
        n = _cross(s0, s1)
        l2 = f32(np.dot(n, n))
        if l2 > 0:
            n = n * (f32(1.0) / np.sqrt(l2, dtype=f32))
            side1 = [s0, v[2] - v[1], v[0] - v[2]]
            side2 = [s1, v[0] - v[1], v[1] - v[2]]
            for j in range(3):
                ang = _unit_angle(_normalize(side1[j]), _normalize(side2[j]))
                normals[f[j]] = normals[f[j]] + n * f32(ang)
    for i in range(len(normals)):
        ln = np.sqrt(np.dot(normals[i], normals[i]).astype(f32), dtype=f32)
        normals[i] = normals[i] / ln if ln != 0 else np.array([1, 0, 0], dtype=f32)
    return normals.astype(f32)


# (mesh file, bsdf id) in XML document order: resources/data/scenes/cbox/fragments/shapes.xml
CBOX_SHAPES = [
    ("cbox_luminaire.obj", "light", (0.0, -0.5, 0.0)),
    ("cbox_floor.obj", "white", (0, 0, 0)),
    ("cbox_ceiling.obj", "white", (0, 0, 0)),
    ("cbox_back.obj", "white", (0, 0, 0)),
    ("cbox_greenwall.obj", "green", (0, 0, 0)),
    ("cbox_redwall.obj", "red", (0, 0, 0)),
    ("cbox_smallbox.obj", "box", (0, 0, 0)),
    ("cbox_largebox.obj", "box", (0, 0, 0)),
]


def make_cbox_meshes():
    out = {}
    for i, (fn, bsdf, tr) in enumerate(CBOX_SHAPES):
        p, n, f = load_obj_like_reference(os.path.join(CBOX, "meshes", fn), tr)
        out[f"m{i}_positions"] = p; out[f"m{i}_normals"] = n; out[f"m{i}_faces"] = f
        out[f"m{i}_bsdf"] = np.array(bsdf)
    np.savez_compressed(os.path.join(HERE, "cbox_meshes.npz"), **out)
    print("cbox_meshes.npz:", sum(len(out[f"m{i}_faces"]) for i in range(len(CBOX_SHAPES))), "triangles")


def read_pfm(fn):
    with open(fn, "rb") as f:
        hdr = f.readline().strip()
        w, h = map(int, f.readline().split())
        scale = float(f.readline())
        c = 3 if hdr == b"PF" else 1
        a = np.frombuffer(f.read(), dtype="<f4" if scale < 0 else ">f4").reshape(h, w, c)
    return np.ascontiguousarray(a[::-1]).astype(np.float32)     # PFM rows are bottom-up


CBOX_XML = """<scene version="2.0.0">
    <integrator type="path">
        <integer name="max_depth" value="{max_depth}"/>
        <integer name="rr_depth" value="{rr_depth}"/>
        <integer name="block_size" value="{block_size}"/>
        {spp_pass}
    </integrator>
    <sensor type="perspective">
        <string name="fov_axis" value="smaller"/>
        <float name="near_clip" value="10"/>
        <float name="far_clip" value="2800"/>
        <float name="focus_distance" value="1000"/>
        <float name="fov" value="39.3077"/>
        <transform name="to_world">
            <lookat origin="278, 273, -800" target="278, 273, -799" up="0, 1, 0"/>
        </transform>
        <sampler type="independent">
            <integer name="sample_count" value="{spp}"/>
            <integer name="seed" value="{seed}"/>
        </sampler>
        <film type="hdrfilm">
            <integer name="width" value="{width}"/>
            <integer name="height" value="{height}"/>
            {crop}
            {rfilter}
            <string name="pixel_format" value="rgb"/>
            <string name="component_format" value="float32"/>
            <string name="file_format" value="pfm"/>
        </film>
    </sensor>
    {bsdfs}
    {shapes}
</scene>
"""

CBOX_BSDFS = {
    "box": (0.45, 0.30, 0.90), "white": (0.885809, 0.698859, 0.666422),
    "red": (0.570068, 0.0430135, 0.0443706), "green": (0.105421, 0.37798, 0.076425),
    "light": (0.936461, 0.740433, 0.705267),
}
CBOX_RADIANCE = (18.387, 10.9873, 2.75357)


ENV_TO_WORLD = [0, 0, 1, 0,  0, 1, 0, 0,  -1, 0, 0, 0,  0, 0, 0, 1]      # exact 90-degree rotation about y
# which cbox meshes are dropped for the environment-lit variants
ENV_DROPPED = {
    "only": ("cbox_luminaire.obj", "cbox_ceiling.obj", "cbox_back.obj", "cbox_greenwall.obj", "cbox_redwall.obj"),
    "with_area": ("cbox_ceiling.obj", "cbox_back.obj", "cbox_greenwall.obj"),
    "constant": ("cbox_luminaire.obj", "cbox_ceiling.obj", "cbox_back.obj"),      # constant background + nothing else
    "constant_area": ("cbox_ceiling.obj", "cbox_greenwall.obj"),
}
CONSTANT_RADIANCE = (0.8, 1.1, 1.6)
# spot light looking down from under the ceiling (exact-in-fp32 to_world: local z -> world -y); directional light along a
# tilted axis (ORACLE ONLY in round 1)
SPOT_LIGHT = dict(to_world=[[1, 0, 0, 278], [0, 0, -1, 540], [0, 1, 0, 279.5], [0, 0, 0, 1]], intensity=(900000.0, 700000.0, 500000.0),
                  cutoff_angle=35.0, beam_width=20.0)
DIRECTIONAL_LIGHT = dict(to_world=[[0.8, 0, 0.6, 0], [0, 1, 0, 0], [-0.6, 0, 0.8, 0], [0, 0, 0, 1]], irradiance=(3.0, 2.5, 2.0))
POINT_LIGHT = dict(position=(278.0, 500.0, 279.5), intensity=(800000.0, 600000.0, 400000.0))


def env_image(w=48, h=24):
    """Deterministic lat-long environment: sky gradient + a bright warm patch + a cool patch."""
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((h, w, 3), np.float32)
    img[..., 0] = 0.3 + 0.5 * (y / (h - 1))
    img[..., 1] = 0.4 + 0.3 * np.cos(x / (w - 1) * np.float32(6.2831853)) ** 2
    img[..., 2] = 0.9 - 0.6 * (y / (h - 1))
    img[4:8, 10:15] += np.array([30.0, 24.0, 12.0], np.float32)
    img[9:12, 30:36] += np.array([2.0, 5.0, 9.0], np.float32)
    return img.astype(np.float32)


def write_pfm(path, img):
    h, w = img.shape[:2]
    with open(path, "wb") as f:
        f.write((b"PF" if img.ndim == 3 else b"Pf") + b"\n%d %d\n-1.0\n" % (w, h))
        f.write(np.ascontiguousarray(img[::-1], "<f4").tobytes())      # PFM rows are bottom-up


def tex_images():
    """Deterministic small bitmaps in [0, 1]: an 8x6 RGB one and a 5x7 single-channel one."""
    y, x = np.mgrid[0:6, 0:8].astype(np.float32)
    rgb = np.stack([0.15 + 0.1 * x, 0.9 - 0.12 * y, 0.2 + 0.05 * ((x * 3 + y * 5) % 7)], -1).astype(np.float32)
    y, x = np.mgrid[0:7, 0:5].astype(np.float32)
    gray = (0.1 + 0.125 * ((x * 2 + y * 3) % 7)).astype(np.float32)
    return np.clip(rgb, 0, 1), np.clip(gray, 0, 1)


# to_uv transforms with exactly representable entries: scale then translate (xml.cpp composes left-multiplying)
TEX_SPEC = {
    # bsdf id: (plugin, texture kind, kwargs for SceneBuffers.add_*, xml of the texture)
    "white": ("diffuse", "checkerboard", dict(color0=(0.8, 0.7, 0.6), color1=(0.2, 0.25, 0.3), scale=(4.0, 4.0), translate=(0.0, 0.0))),
    "green": ("diffuse", "bitmap_rgb", dict(filter_type="bilinear", wrap_mode="repeat", scale=(2.0, 3.0), translate=(0.25, 0.5))),
    "red": ("diffuse", "bitmap_gray", dict(filter_type="nearest", wrap_mode="mirror", scale=(3.0, 2.0), translate=(-0.5, 0.25))),
    "box": ("plastic", "bitmap_rgb", dict(filter_type="bilinear", wrap_mode="clamp", scale=(1.5, 1.5), translate=(-0.25, -0.125))),
}


def tex_bsdf_xml(bsdf_id, mesh_dir):
    plugin, kind, kw = TEX_SPEC[bsdf_id]
    sx, sy = kw["scale"]; tx, ty = kw["translate"]
    xf = f'<transform name="to_uv"><scale x="{sx}" y="{sy}"/><translate x="{tx}" y="{ty}"/></transform>'
    pname = "reflectance" if plugin == "diffuse" else "diffuse_reflectance"
    if kind == "checkerboard":
        c0, c1 = kw["color0"], kw["color1"]
        tex = (f'<texture type="checkerboard" name="{pname}"><rgb name="color0" value="{c0[0]}, {c0[1]}, {c0[2]}"/>'
               f'<rgb name="color1" value="{c1[0]}, {c1[1]}, {c1[2]}"/>{xf}</texture>')
    else:
        fn = os.path.join(mesh_dir, "tex_rgb.pfm" if kind == "bitmap_rgb" else "tex_gray.pfm")
        tex = (f'<texture type="bitmap" name="{pname}"><string name="filename" value="{fn}"/>'
               f'<string name="filter_type" value="{kw["filter_type"]}"/><string name="wrap_mode" value="{kw["wrap_mode"]}"/>'
               f'<boolean name="raw" value="true"/>{xf}</texture>')
    extra = '<float name="int_ior" value="1.6"/>' if plugin == "plastic" else ""
    return f'<bsdf type="{plugin}" id="{bsdf_id}">{tex}{extra}</bsdf>'


def write_uv_obj(src, dst):
    """Copy an OBJ adding one vt per vertex: a generic non-degenerate projection of the position."""
    out, n = [], 0
    for line in open(src):
        cur = line.strip()
        if cur.startswith("v "):
            x, y, z = [float(t) for t in cur.split()[1:4]]
            out.append(line if line.endswith("\n") else line + "\n")
            out.append("vt %r %r\n" % (float(f32(x / 500.0 + y / 700.0)), float(f32(z / 500.0 + y / 900.0))))
            n += 1
        elif cur.startswith("f "):
            toks = []
            for t in cur.split()[1:]:
                p = t.split("/")
                toks.append(f"{p[0]}/{p[0]}" + (f"/{p[2]}" if len(p) > 2 and p[2] else ""))
            out.append("f " + " ".join(toks) + "\n")
        else:
            out.append(line)
    # all vt after all v is not required by the format, but keep the indices aligned: vt k belongs to v k
    open(dst, "w").writelines(out)


# analytic shapes variant: the luminaire mesh becomes a rectangle emitter (exactly invertible to_world: the
# [-1,1]^2 square -> 128 x 64 patch facing down), the small box becomes a diffuse sphere, plus a small
# emissive sphere (sphere emitter sampling) and a rectangle standing on the floor with flipped normals
SHAPES_SPEC = dict(
    rect_light=dict(to_world=[[64, 0, 0, 278], [0, 0, -1, 548], [0, 32, 0, 280], [0, 0, 0, 1]], radiance=(18.387, 10.9873, 2.75357)),
    sphere=dict(center=(186.0, 64.0, 169.0), radius=64.0, bsdf="box"),
    sphere_light=dict(center=(420.0, 360.0, 300.0), radius=16.0, radiance=(30.0, 30.0, 60.0)),
    rect_flip=dict(to_world=[[32, 0, 0, 100], [0, 64, 0, 64], [0, 0, 1, 400], [0, 0, 0, 1]], bsdf="green", flip_normals=True),
)


def shapes_xml(textured=False, mesh_dir=None):
    sp = SHAPES_SPEC
    if textured:
        # same shapes, but the sphere carries a checkerboard and the free-standing rectangle a bitmap: exercises the
        # analytic uv parameterisations (sphere.cpp:352-361, rectangle.cpp:196-197)
        rgb, _ = tex_images()
        write_pfm(os.path.join(mesh_dir, "tex_rgb.pfm"), rgb)
    mat = lambda M: " ".join(repr(float(x)) for row in M for x in row)
    rl, s1, sl, rf = sp["rect_light"], sp["sphere"], sp["sphere_light"], sp["rect_flip"]
    return [
        f'<shape type="rectangle"><transform name="to_world"><matrix value="{mat(rl["to_world"])}"/></transform><ref id="light"/>'
        f'<emitter type="area"><rgb name="radiance" value="{rl["radiance"][0]}, {rl["radiance"][1]}, {rl["radiance"][2]}"/></emitter></shape>',
        f'<shape type="sphere"><point name="center" x="{s1["center"][0]}" y="{s1["center"][1]}" z="{s1["center"][2]}"/>'
        f'<float name="radius" value="{s1["radius"]}"/>' + (
            '<bsdf type="diffuse"><texture type="checkerboard" name="reflectance"><rgb name="color0" value="0.9, 0.8, 0.2"/>'
            '<rgb name="color1" value="0.1, 0.2, 0.6"/><transform name="to_uv"><scale x="8" y="4"/></transform></texture></bsdf>'
            if textured else f'<ref id="{s1["bsdf"]}"/>') + '</shape>',
        f'<shape type="sphere"><point name="center" x="{sl["center"][0]}" y="{sl["center"][1]}" z="{sl["center"][2]}"/>'
        f'<float name="radius" value="{sl["radius"]}"/><ref id="light"/>'
        f'<emitter type="area"><rgb name="radiance" value="{sl["radiance"][0]}, {sl["radiance"][1]}, {sl["radiance"][2]}"/></emitter></shape>',
        f'<shape type="rectangle"><transform name="to_world"><matrix value="{mat(rf["to_world"])}"/></transform>'
        f'<boolean name="flip_normals" value="true"/>' + (
            f'<bsdf type="twosided"><bsdf type="diffuse"><texture type="bitmap" name="reflectance"><string name="filename" value="{mesh_dir}/tex_rgb.pfm"/>'
            '<boolean name="raw" value="true"/><transform name="to_uv"><scale x="2" y="2"/></transform></texture></bsdf></bsdf>'
            if textured else f'<ref id="{rf["bsdf"]}"/>') + '</shape>',
    ]


EXTRA_EMITTER_OBJ = """v 150 420 150
v 260 420 150
v 260 420 270
v 150 420 270
f 1 2 3 4
"""
EXTRA_EMITTER_RADIANCE = (4.0, 16.0, 6.0)


def cbox_xml(width, height, spp, max_depth=6, rr_depth=5, block_size=32, samples_per_pass=None,
             rfilter="box", seed=0, materials=None, no_patch=False, mesh_dir=None, lift=0.0, extra_emitter=False, instances=None, envmap=None,
             textures=False, crop=None, analytic=False, integrator="path", sampler="independent", jitter=True, point_light=None, direct=None):
    """materials: optional {bsdf_id: xml-string} overriding the diffuse definitions. direct: (emitter_samples, bsdf_samples)"""
    bs = []
    for k, v in CBOX_BSDFS.items():
        if textures and k in TEX_SPEC:
            bs.append(tex_bsdf_xml(k, mesh_dir))
        elif materials and k in materials:
            bs.append(materials[k].replace("<bsdf ", f'<bsdf id="{k}" ', 1))
        else:
            bs.append(f'<bsdf type="diffuse" id="{k}"><rgb name="reflectance" value="{v[0]}, {v[1]}, {v[2]}"/></bsdf>')
    shapes = []
    for fn, bsdf, tr in CBOX_SHAPES:
        if point_light in ("only", "spot") and fn == "cbox_luminaire.obj":
            continue                       # lit by the point light alone
        if analytic and fn in ("cbox_luminaire.obj", "cbox_smallbox.obj"):
            continue                       # replaced by analytic shapes (appended below)
        if envmap is not None and fn in ENV_DROPPED[envmap]:
            continue                       # open scene: the environment must be visible
        if instances is not None and fn == "cbox_largebox.obj":
            continue
        if instances is not None and fn == "cbox_smallbox.obj":
            # shapegroup + instances (src/librender/shapegroup.cpp, src/shapes/instance.cpp)
            shapes.append(f'<shape type="shapegroup" id="grp"><shape type="obj"><string name="filename" value="{CBOX}/meshes/{fn}"/>'
                          f'<ref id="{bsdf}"/></shape></shape>')
            for m in instances:
                vals = " ".join(repr(float(x)) for row in m for x in row)
                shapes.append(f'<shape type="instance"><ref id="grp"/><transform name="to_world"><matrix value="{vals}"/></transform></shape>')
            continue
        if lift and bsdf == "box":
            tr = (0.0, float(lift), 0.0)     # float the two boxes: no face coplanar with the floor
        t = f'<transform name="to_world"><translate x="{tr[0]}" y="{tr[1]}" z="{tr[2]}"/></transform>' if any(tr) else ""
        em = ('<emitter type="area"><rgb name="radiance" value="18.387, 10.9873, 2.75357"/></emitter>'
              if bsdf == "light" else "")
        path = f"{CBOX}/meshes/{fn}"
        if no_patch and fn == "cbox_ceiling.obj":
            # the shipped ceiling has a second quad (f 8 7 6 5) coplanar with the first: z-fighting
            # whose winner depends on last-ulp rounding (DESIGN.md "Float parity"). Drop it.
            path = os.path.join(mesh_dir, fn)
            lines = [l for l in open(f"{CBOX}/meshes/{fn}") if l.strip() != "f 8 7 6 5"]
            open(path, "w").writelines(lines)
        if textures:           # every mesh gets texture coordinates (=> uv-derived dp_du on all surfaces)
            dst = os.path.join(mesh_dir, "uv_" + fn)
            write_uv_obj(path, dst)
            path = dst
        shapes.append(f'<shape type="obj"><string name="filename" value="{path}"/>{t}<ref id="{bsdf}"/>{em}</shape>')
    mat = lambda M: " ".join(repr(float(x)) for row in M for x in row)
    if point_light == "spot":
        sl = SPOT_LIGHT
        shapes.append(f'<emitter type="spot"><transform name="to_world"><matrix value="{mat(sl["to_world"])}"/></transform>'
                      f'<rgb name="intensity" value="{sl["intensity"][0]}, {sl["intensity"][1]}, {sl["intensity"][2]}"/>'
                      f'<float name="cutoff_angle" value="{sl["cutoff_angle"]}"/><float name="beam_width" value="{sl["beam_width"]}"/></emitter>')
    elif point_light == "directional":
        dl = DIRECTIONAL_LIGHT
        shapes.append(f'<emitter type="directional"><transform name="to_world"><matrix value="{mat(dl["to_world"])}"/></transform>'
                      f'<rgb name="irradiance" value="{dl["irradiance"][0]}, {dl["irradiance"][1]}, {dl["irradiance"][2]}"/></emitter>')
    elif point_light:                  # PointLight (src/emitters/point.cpp), after the shapes
        p, i = POINT_LIGHT["position"], POINT_LIGHT["intensity"]
        shapes.append(f'<emitter type="point"><point name="position" x="{p[0]}" y="{p[1]}" z="{p[2]}"/>'
                      f'<rgb name="intensity" value="{i[0]}, {i[1]}, {i[2]}"/></emitter>')
    if analytic:
        shapes += shapes_xml(textured=analytic == "textured", mesh_dir=mesh_dir)
    if textures:
        rgb, gray = tex_images()
        write_pfm(os.path.join(mesh_dir, "tex_rgb.pfm"), rgb); write_pfm(os.path.join(mesh_dir, "tex_gray.pfm"), gray)
    if extra_emitter:
        # a second area emitter (scene.cpp:176-198: uniform emitter selection): a small quad facing down
        path = os.path.join(mesh_dir, "extra_emitter.obj")
        open(path, "w").write(EXTRA_EMITTER_OBJ)
        shapes.append(f'<shape type="obj"><string name="filename" value="{path}"/><ref id="white"/>'
                      '<emitter type="area"><rgb name="radiance" value="4.0, 16.0, 6.0"/></emitter></shape>')
    if envmap is not None and envmap.startswith("constant"):       # ConstantBackgroundEmitter (src/emitters/constant.cpp)
        shapes.append(f'<emitter type="constant"><rgb name="radiance" value="{CONSTANT_RADIANCE[0]}, {CONSTANT_RADIANCE[1]}, {CONSTANT_RADIANCE[2]}"/></emitter>')
    elif envmap is not None:
        path = os.path.join(mesh_dir, "env.pfm")
        write_pfm(path, env_image())
        shapes.append(f'<emitter type="envmap"><string name="filename" value="{path}"/><float name="scale" value="1.5"/>'
                      f'<transform name="to_world"><matrix value="{" ".join(str(float(x)) for x in ENV_TO_WORLD)}"/></transform></emitter>')
    rf = '<rfilter type="box"/>' if rfilter == "box" else '<rfilter type="gaussian"/>'
    sp = f'<integer name="samples_per_pass" value="{samples_per_pass}"/>' if samples_per_pass else ""
    cr = ""
    if crop:          # film.cpp:17-27
        cr = (f'<integer name="crop_offset_x" value="{crop[0]}"/><integer name="crop_offset_y" value="{crop[1]}"/>'
              f'<integer name="crop_width" value="{crop[2]}"/><integer name="crop_height" value="{crop[3]}"/>')
    xml = CBOX_XML.format(crop=cr, width=width, height=height, spp=spp, max_depth=max_depth, rr_depth=rr_depth,
                          block_size=block_size, spp_pass=sp, rfilter=rf, seed=seed,
                          bsdfs="\n    ".join(bs), shapes="\n    ".join(shapes))
    if sampler != "independent":        # src/samplers/stratified.cpp
        xml = xml.replace('<sampler type="independent">', f'<sampler type="{sampler}">' + ("" if jitter else '<boolean name="jitter" value="false"/>'))
    if integrator == "direct":         # DirectIntegrator (src/integrators/direct.cpp): no depth parameters
        xml = xml.replace('<integrator type="path">', '<integrator type="direct">' + (
            f'<integer name="emitter_samples" value="{direct[0]}"/><integer name="bsdf_samples" value="{direct[1]}"/>' if direct else ""))
        xml = re.sub(r'\s*<integer name="(max_depth|rr_depth)" value="[^"]*"/>', "", xml)
    return xml


def run_reference(xml_fn, threads=None):
    """xml_fn(mesh_dir) -> scene xml"""
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = DIST + ":" + env.get("LD_LIBRARY_PATH", "")
    with tempfile.TemporaryDirectory() as td:
        sp = os.path.join(td, "scene.xml")
        open(sp, "w").write(xml_fn(td))
        out = os.path.join(td, "out.pfm")
        cmd = [os.path.join(DIST, "mitsuba"), "-m", "scalar_rgb", "-o", out, sp]
        if threads:
            cmd[1:1] = ["-t", str(threads)]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True)
        if r.returncode != 0 or not os.path.exists(out):
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("reference render failed")
        return read_pfm(out), r.stdout + r.stderr


DIELECTRIC = '<bsdf type="dielectric"><float name="int_ior" value="1.5"/><float name="ext_ior" value="1.0"/></bsdf>'
ROUGH_GGX = '<bsdf type="roughconductor"><string name="distribution" value="ggx"/><float name="alpha" value="0.1"/></bsdf>'
ROUGH_BECK = '<bsdf type="roughconductor"><float name="alpha" value="0.2"/></bsdf>'
CONDUCTOR = '<bsdf type="conductor"><rgb name="eta" value="0.2, 0.92, 1.1"/><rgb name="k" value="3.9, 2.45, 2.14"/></bsdf>'
MIRROR = '<bsdf type="conductor"><rgb name="specular_reflectance" value="0.9, 0.8, 0.7"/></bsdf>'     # material "none"
PLASTIC = '<bsdf type="plastic"><rgb name="diffuse_reflectance" value="0.1, 0.27, 0.36"/><float name="int_ior" value="1.9"/></bsdf>'
PLASTIC_SPEC = ('<bsdf type="plastic"><rgb name="diffuse_reflectance" value="0.7, 0.6, 0.5"/>'
                '<rgb name="specular_reflectance" value="0.8, 0.9, 1.0"/><float name="int_ior" value="1.49"/>'
                '<float name="ext_ior" value="1.000277"/></bsdf>')
ROUGH_GLASS_GGX = ('<bsdf type="roughdielectric"><string name="distribution" value="ggx"/><float name="alpha" value="0.1"/>'
                   '<float name="int_ior" value="1.5"/><float name="ext_ior" value="1.0"/></bsdf>')
ROUGH_GLASS_BECK = ('<bsdf type="roughdielectric"><float name="alpha" value="0.25"/><float name="int_ior" value="1.33"/>'
                    '<float name="ext_ior" value="1.0"/><rgb name="specular_transmittance" value="0.9, 0.95, 1.0"/></bsdf>')
TWOSIDED_1 = '<bsdf type="twosided"><bsdf type="diffuse"><rgb name="reflectance" value="0.45, 0.3, 0.9"/></bsdf></bsdf>'
TWOSIDED_2 = ('<bsdf type="twosided"><bsdf type="diffuse"><rgb name="reflectance" value="0.1, 0.6, 0.1"/></bsdf>'
              '<bsdf type="roughconductor"><float name="alpha" value="0.2"/></bsdf></bsdf>')
PLASTIC_NONLINEAR = ('<bsdf type="plastic"><rgb name="diffuse_reflectance" value="0.8, 0.3, 0.2"/><boolean name="nonlinear" value="true"/></bsdf>')

# exact-in-fp32 instance transforms: identity, half-size copy, 90-degree rotation about y (floating),
# mirrored copy (floating)
INSTANCES = [
    [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]],
    [[0.5, 0, 0, 330], [0, 0.5, 0, 0], [0, 0, 0.5, 250], [0, 0, 0, 1]],
    [[0, 0, 1, 250], [0, 1, 0, 200], [-1, 0, 0, 560], [0, 0, 0, 1]],
    [[-1, 0, 0, 420], [0, 1, 0, 300], [0, 0, 1, 0], [0, 0, 0, 1]],
]

GOLDEN = {
    # name: kwargs for cbox_xml
    "cbox_64_spp16": dict(width=64, height=64, spp=16),
    "cbox_256_spp16": dict(width=256, height=256, spp=16),                    # BASELINE configs[0] (C1)
    "cbox_nopatch_256_spp16": dict(width=256, height=256, spp=16, no_patch=True),
    "cbox_nopatch_64_spp64": dict(width=64, height=64, spp=64, no_patch=True),
    "cbox_64_spp16_pass1": dict(width=64, height=64, spp=16, samples_per_pass=1),
    "cbox_64_spp16_pass4": dict(width=64, height=64, spp=16, samples_per_pass=4),
    "cbox_80x48_spp8": dict(width=80, height=48, spp=8),                      # ragged blocks
    "cbox_64_spp16_gauss": dict(width=64, height=64, spp=16, rfilter="gaussian"),
    "cbox_64_spp16_depth3_rr2": dict(width=64, height=64, spp=16, max_depth=3, rr_depth=2),
    "cbox_64_spp16_bs16_seed7": dict(width=64, height=64, spp=16, block_size=16, seed=7),
    "cbox_64_spp16_dielectric": dict(width=64, height=64, spp=16, materials={"box": DIELECTRIC}),
    "cbox_64_spp16_dielectric_lift": dict(width=64, height=64, spp=16, materials={"box": DIELECTRIC}, lift=10.0),
    "cbox_64_spp16_two_emitters": dict(width=64, height=64, spp=16, no_patch=True, extra_emitter=True),
    "cbox_64_spp16_instances": dict(width=64, height=64, spp=16, no_patch=True, instances=INSTANCES),
    "cbox_64_spp16_env": dict(width=64, height=64, spp=16, envmap="only"),
    "cbox_64_spp16_env_area": dict(width=64, height=64, spp=16, envmap="with_area"),
    # BASELINE configs[2] in miniature: rough conductor (GGX) lit by an environment map
    "cbox_64_spp16_env_ggx": dict(width=64, height=64, spp=16, envmap="only", materials={"box": ROUGH_GGX}),
    "cbox_64_spp16_ggx": dict(width=64, height=64, spp=16, materials={"box": ROUGH_GGX}),
    "cbox_64_spp16_beckmann": dict(width=64, height=64, spp=16, materials={"box": ROUGH_BECK}),
    # smooth conductor / plastic
    "cbox_64_spp16_conductor": dict(width=64, height=64, spp=16, no_patch=True, materials={"box": CONDUCTOR, "green": MIRROR}),
    # film crop window, unbounded depth (Russian roulette only), emitter-only / single-bounce depths, 1 spp
    "cbox_64_spp16_crop": dict(width=64, height=64, spp=16, no_patch=True, crop=(8, 16, 40, 24)),
    "cbox_64_spp8_depthinf_rr3": dict(width=64, height=64, spp=8, no_patch=True, max_depth=-1, rr_depth=3),
    "cbox_64_spp16_depth1": dict(width=64, height=64, spp=16, no_patch=True, max_depth=1),
    "cbox_64_spp16_depth2": dict(width=64, height=64, spp=16, no_patch=True, max_depth=2),
    "cbox_48_spp1": dict(width=48, height=48, spp=1, no_patch=True),
    # analytic shapes: rectangle (emitter + flipped), sphere (diffuse + emitter)
    "cbox_64_spp16_shapes": dict(width=64, height=64, spp=16, no_patch=True, analytic=True),
    # twosided: the free-standing flipped rectangle ("green": diffuse front / rough conductor back) is seen from both sides
    "cbox_64_spp16_twosided": dict(width=64, height=64, spp=16, no_patch=True, analytic=True, materials={"green": TWOSIDED_2, "box": TWOSIDED_1}),
    # the reference's `direct` integrator with its default shading_samples = 1 draws the same numbers in the same order as
    # `path` with max_depth = 2 (direct.cpp:106-196 vs path.cpp:124-208): rendered by `direct`, loaded as max_depth 2
    # ... and with emitter_samples = 2, bsdf_samples = 3 (direct.cpp:106-196 proper)
    "cbox_64_spp8_direct_e2_b3": dict(width=64, height=64, spp=8, no_patch=True, analytic=True, integrator="direct", direct=(2, 3)),
    "cbox_64_spp8_direct_e3_b0": dict(width=64, height=64, spp=8, no_patch=True, integrator="direct", direct=(3, 0), materials=dict(box=ROUGH_GGX)),
    "cbox_64_spp8_direct_e0_b2": dict(width=64, height=64, spp=8, no_patch=True, integrator="direct", direct=(0, 2), materials=dict(box=DIELECTRIC), lift=5.0),
    "cbox_64_spp16_direct": dict(width=64, height=64, spp=16, no_patch=True, analytic=True, integrator="direct", max_depth=2),
    # stratified sampler (with and without jitter; several passes re-seed and restart the strata like the reference)
    "cbox_64_spp16_stratified": dict(width=64, height=64, spp=16, no_patch=True, sampler="stratified"),
    "cbox_64_spp16_stratified_nojitter_pass4": dict(width=64, height=64, spp=16, no_patch=True, sampler="stratified", jitter=False, samples_per_pass=4),
    "cbox_64_spp16_multijitter": dict(width=64, height=64, spp=16, no_patch=True, sampler="multijitter"),
    "cbox_64_spp8_multijitter_nojitter": dict(width=64, height=64, spp=8, no_patch=True, sampler="multijitter", jitter=False),   # 4 x 2 strata
    "cbox_64_spp16_shapes_textured": dict(width=64, height=64, spp=16, no_patch=True, analytic="textured"),
    "cbox_64_spp16_ldsampler": dict(width=64, height=64, spp=16, no_patch=True, sampler="ldsampler"),
    # orthogonal-array sampler (Bose construction, strength 2): spp = p^2 with p prime
    "cbox_64_spp25_orthogonal": dict(width=64, height=64, spp=25, no_patch=True, sampler="orthogonal"),
    "cbox_64_spp9_orthogonal_nojitter": dict(width=64, height=64, spp=9, no_patch=True, sampler="orthogonal", jitter=False),
    "cbox_64_spp16_constant": dict(width=64, height=64, spp=16, envmap="constant"),
    "cbox_64_spp16_constant_area": dict(width=64, height=64, spp=16, envmap="constant_area"),
    "cbox_64_spp16_point": dict(width=64, height=64, spp=16, no_patch=True, point_light="only"),
    "cbox_64_spp16_point_area": dict(width=64, height=64, spp=16, no_patch=True, point_light="with_area"),
    # rough dielectric (ORACLE ONLY in round 1; the CUDA path rejects it): boxes lifted off the floor like the smooth dielectric case
    "cbox_64_spp16_roughglass_ggx": dict(width=64, height=64, spp=16, no_patch=True, lift=10.0, materials={"box": ROUGH_GLASS_GGX}),
    "cbox_64_spp16_roughglass_beckmann": dict(width=64, height=64, spp=16, no_patch=True, lift=10.0, materials={"box": ROUGH_GLASS_BECK}),
    "cbox_64_spp16_spot": dict(width=64, height=64, spp=16, no_patch=True, point_light="spot"),
    "cbox_64_spp16_directional": dict(width=64, height=64, spp=16, envmap="constant_area", point_light="directional"),   # open box: + constant + area
    "cbox_64_spp16_textures": dict(width=64, height=64, spp=16, no_patch=True, textures=True),
    "cbox_64_spp16_plastic": dict(width=64, height=64, spp=16, no_patch=True, materials={"box": PLASTIC, "white": PLASTIC_SPEC, "red": PLASTIC_NONLINEAR}),
}


def write_obj(path, pos, nrm, faces):
    with open(path, "w") as f:
        for p in pos:
            f.write("v %s %s %s\n" % tuple(repr(float(x)) for x in p))
        for n in nrm:
            f.write("vn %s %s %s\n" % tuple(repr(float(x)) for x in n))
        for a, b, c in faces:
            f.write("f %d//%d %d//%d %d//%d\n" % (a + 1, a + 1, b + 1, b + 1, c + 1, c + 1))


def veach_xml(mesh_dir, width, height, spp):
    sys.path.insert(0, ROOT)
    from mitsuba2_b200 import scenes
    shapes = []
    for m in scenes.veach_mis_meshes():
        pos, nrm, faces = m["mesh"]
        path = os.path.join(mesh_dir, m["name"] + ".obj")
        write_obj(path, pos, nrm, faces)
        mat = m["material"]
        if mat[0] == "diffuse":
            b = f'<bsdf type="diffuse"><rgb name="reflectance" value="{mat[1][0]}, {mat[1][1]}, {mat[1][2]}"/></bsdf>'
        elif mat[0] == "dielectric":
            b = f'<bsdf type="dielectric"><float name="int_ior" value="{mat[1]}"/><float name="ext_ior" value="1.0"/></bsdf>'
        else:
            b = f'<bsdf type="roughconductor"><string name="distribution" value="{mat[1]}"/><float name="alpha" value="{mat[2]}"/></bsdf>'
        em = ""
        if "radiance" in m:
            r = m["radiance"]
            em = f'<emitter type="area"><rgb name="radiance" value="{r[0]}, {r[1]}, {r[2]}"/></emitter>'
        shapes.append(f'<shape type="obj"><string name="filename" value="{path}"/>{b}{em}</shape>')
    return f"""<scene version="2.0.0">
    <integrator type="path"><integer name="max_depth" value="6"/><integer name="block_size" value="32"/></integrator>
    <sensor type="perspective"><string name="fov_axis" value="x"/><float name="near_clip" value="0.1"/><float name="far_clip" value="100"/>
        <float name="fov" value="28"/>
        <transform name="to_world"><lookat origin="0, 2, 15" target="0, -2, 2.5" up="0, 1, 0"/></transform>
        <sampler type="independent"><integer name="sample_count" value="{spp}"/></sampler>
        <film type="hdrfilm"><integer name="width" value="{width}"/><integer name="height" value="{height}"/><rfilter type="box"/>
            <string name="pixel_format" value="rgb"/><string name="component_format" value="float32"/><string name="file_format" value="pfm"/></film>
    </sensor>
    {"".join(shapes)}
</scene>"""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    if not os.path.isdir(REF) or not os.path.exists(os.path.join(DIST, "mitsuba")):
        sys.exit("needs /root/reference and oracle/_ref/dist/mitsuba (run oracle/build_ref.sh)")
    if a.only in (None, "meshes"):
        make_cbox_meshes()
    info = open(os.path.join(ROOT, "oracle", "_ref", "BUILD_INFO.txt")).read().strip()
    if a.only in (None, "veach_96x54_spp16"):
        img, log = run_reference(lambda td: veach_xml(td, 96, 54, 16))
        np.savez_compressed(os.path.join(HERE, "ref_veach_96x54_spp16.npz"), rgb=img,
                            params=np.array(repr(dict(scene="veach", width=96, height=54, spp=16))), oracle_build=np.array(info))
        print("ref_veach_96x54_spp16.npz", img.shape, "mean", img.mean(axis=(0, 1)))
    for name, kw in GOLDEN.items():
        if a.only and a.only != name:
            continue
        img, log = run_reference(lambda td: cbox_xml(mesh_dir=td, **kw))
        kw_s = {k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items()}) for k, v in kw.items()}
        extra = {}
        if kw.get("extra_emitter"):
            with tempfile.TemporaryDirectory() as td:
                p = os.path.join(td, "e.obj"); open(p, "w").write(EXTRA_EMITTER_OBJ)
                ep, en, ef = load_obj_like_reference(p)
            extra = dict(extra_positions=ep, extra_normals=en, extra_faces=ef, extra_radiance=np.array(EXTRA_EMITTER_RADIANCE, np.float32))
        if kw.get("envmap", "").startswith("constant"):
            extra.update(const_radiance=np.array(CONSTANT_RADIANCE, np.float32),
                         env_dropped=np.array([CBOX_SHAPES.index(next(s for s in CBOX_SHAPES if s[0] == fn)) for fn in ENV_DROPPED[kw["envmap"]]]))
        elif kw.get("envmap"):
            extra.update(env_rgb=env_image(), env_to_world=np.array(ENV_TO_WORLD, np.float32), env_scale=np.float32(1.5),
                         env_dropped=np.array([CBOX_SHAPES.index(next(s for s in CBOX_SHAPES if s[0] == fn)) for fn in ENV_DROPPED[kw["envmap"]]]))
        if kw.get("point_light"):
            extra.update(point_light=np.array(repr(dict(spot=SPOT_LIGHT, directional=DIRECTIONAL_LIGHT).get(kw["point_light"], POINT_LIGHT))))
        if kw.get("analytic"):
            extra.update(shapes_spec=np.array(repr(SHAPES_SPEC)))
            if kw["analytic"] == "textured":
                extra.update(tex_rgb=tex_images()[0])
        if kw.get("textures"):
            rgb_t, gray_t = tex_images()
            extra.update(tex_rgb=rgb_t, tex_gray=gray_t, tex_spec=np.array(repr(TEX_SPEC)))
            with tempfile.TemporaryDirectory() as td:
                for i, (fn, _, tr) in enumerate(CBOX_SHAPES):
                    src = os.path.join(CBOX, "meshes", fn)
                    if kw.get("no_patch") and fn == "cbox_ceiling.obj":
                        src2 = os.path.join(td, "np_" + fn)
                        open(src2, "w").writelines([l for l in open(src) if l.strip() != "f 8 7 6 5"]); src = src2
                    dst = os.path.join(td, "uv_" + fn); write_uv_obj(src, dst)
                    extra[f"uv_m{i}"] = load_obj_like_reference(dst, tr)[3]
        np.savez_compressed(os.path.join(HERE, f"ref_{name}.npz"), rgb=img, params=np.array(repr(kw_s)),
                            oracle_build=np.array(info), **extra)
        print(f"ref_{name}.npz", img.shape, "mean", img.mean(axis=(0, 1)))


if __name__ == "__main__":
    main()
